//! Safe Rust face of `libtrueno_rag_b200.so`.
//!
//! The reference keeps its public types and signatures; their bodies delegate to this crate:
//!
//! | reference item (file:line)                          | delegates to                         |
//! |------------------------------------------------------|--------------------------------------|
//! | `VectorStore::insert / insert_batch` (index.rs:359)  | [`DenseStore::insert`], [`DenseStore::insert_batch`] |
//! | `VectorStore::search` (index.rs:386)                 | [`DenseStore::search`]               |
//! | `VectorStore::remove` (index.rs:421)                 | [`DenseStore::remove`]               |
//! | `BM25Index::search` (index.rs:212)                   | [`Bm25Device::search`] after [`Bm25Device::freeze`] |
//! | `FusionStrategy::fuse` (fusion.rs:42)                | [`fuse`]                             |
//! | `HybridRetriever::retrieve` (retrieve.rs:175)        | [`hybrid_search`]                    |
//!
//! Chunk ids cross this boundary as `u128` (`ChunkId.0.as_u128()`); the library works on insertion ordinals and
//! returns results in the canonical order *score descending, insertion ordinal ascending* (the reference's order among
//! exact ties is HashMap iteration order, i.e. unspecified).
//!
//! No Rust toolchain exists in the image this repository is built in: the crate is delivered as source and is
//! exercised through the identical C++ mirror (`include/trueno_rag.hpp`), which the test-suite drives.

use std::collections::HashMap;
use std::ffi::CStr;
use std::fmt;
use std::os::raw::c_int;
use std::sync::OnceLock;

use trueno_rag_b200_sys as sys;

/// Errors of the device library, shaped after `trueno_rag::Error` (src/error.rs).
#[derive(Debug, Clone, PartialEq)]
pub enum DeviceError {
    /// -> `Error::InvalidConfig`
    InvalidArg(String),
    /// -> `Error::DimensionMismatch { expected, actual }`
    DimensionMismatch { expected: usize, actual: usize },
    /// -> `Error::VectorStore` (CUDA failure, out of memory, no device: there is no CPU fallback)
    Device(String),
    /// a request outside what the kernels serve (k > 1024, more than 512 query terms, ...)
    Unsupported(String),
}

impl fmt::Display for DeviceError {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result {
        match self {
            DeviceError::InvalidArg(m) => write!(f, "invalid argument: {m}"),
            DeviceError::DimensionMismatch { expected, actual } => {
                write!(f, "dimension mismatch: expected {expected}, got {actual}")
            }
            DeviceError::Device(m) => write!(f, "device error: {m}"),
            DeviceError::Unsupported(m) => write!(f, "unsupported: {m}"),
        }
    }
}

impl std::error::Error for DeviceError {}

pub type Result<T> = std::result::Result<T, DeviceError>;

fn last_error() -> String {
    // SAFETY: trr_last_error returns a NUL-terminated string owned by the library (thread-local, valid until the next call)
    unsafe { CStr::from_ptr(sys::trr_last_error()) }.to_string_lossy().into_owned()
}

fn check(status: c_int) -> Result<()> {
    match status {
        sys::TRR_OK => Ok(()),
        sys::TRR_ERR_INVALID_ARG => Err(DeviceError::InvalidArg(last_error())),
        sys::TRR_ERR_UNSUPPORTED => Err(DeviceError::Unsupported(last_error())),
        _ => Err(DeviceError::Device(last_error())),
    }
}

/// `DistanceMetric` (src/index.rs:310-319)
#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub enum Metric {
    Cosine,
    Euclidean,
    DotProduct,
}

impl Metric {
    fn raw(self) -> c_int {
        match self {
            Metric::Cosine => sys::TRR_METRIC_COSINE,
            Metric::Euclidean => sys::TRR_METRIC_EUCLIDEAN,
            Metric::DotProduct => sys::TRR_METRIC_DOT,
        }
    }
}

/// `FusionStrategy` (src/fusion.rs:9-31) as (kind, parameter)
#[derive(Debug, Clone, Copy, PartialEq)]
pub enum Fusion {
    Rrf { k: f32 },
    Linear { dense_weight: f32 },
    Convex { alpha: f32 },
    Dbsf,
    Union,
    Intersection,
}

impl Fusion {
    fn raw(self) -> (c_int, f32) {
        match self {
            Fusion::Rrf { k } => (sys::TRR_FUSE_RRF, k),
            Fusion::Linear { dense_weight } => (sys::TRR_FUSE_LINEAR, dense_weight),
            Fusion::Convex { alpha } => (sys::TRR_FUSE_CONVEX, alpha),
            Fusion::Dbsf => (sys::TRR_FUSE_DBSF, 0.0),
            Fusion::Union => (sys::TRR_FUSE_UNION, 0.0),
            Fusion::Intersection => (sys::TRR_FUSE_INTERSECTION, 0.0),
        }
    }
}

struct Ctx(*mut sys::trr_ctx);
// SAFETY: every entry point of the library locks the context internally
unsafe impl Send for Ctx {}
unsafe impl Sync for Ctx {}

/// The process-wide device context (device `TRR_DEVICE`, default 0), created on first use.
fn context() -> Result<*mut sys::trr_ctx> {
    static CTX: OnceLock<std::result::Result<Ctx, DeviceError>> = OnceLock::new();
    let r = CTX.get_or_init(|| {
        let device = std::env::var("TRR_DEVICE").ok().and_then(|v| v.parse::<c_int>().ok()).unwrap_or(0);
        let mut ctx: *mut sys::trr_ctx = std::ptr::null_mut();
        // SAFETY: `ctx` is a valid out pointer
        check(unsafe { sys::trr_ctx_create(device, &mut ctx) }).map(|()| Ctx(ctx))
    });
    match r {
        Ok(c) => Ok(c.0),
        Err(e) => Err(e.clone()),
    }
}

// ------------------------------------------------------------------------------------------------
// VectorStore
// ------------------------------------------------------------------------------------------------

/// The embedding slab of a `VectorStore` on the device plus the id <-> ordinal maps.
pub struct DenseStore {
    h: *mut sys::trr_dense,
    dim: usize,
    id_of: Vec<u128>,
    ord_of: HashMap<u128, u32>,
}

// SAFETY: the handle is only used through the library, which serialises access per context
unsafe impl Send for DenseStore {}
unsafe impl Sync for DenseStore {}

impl DenseStore {
    /// `VectorStore::new(config)` (src/index.rs:332-340); `bf16` selects bf16 storage (arithmetic stays f32)
    pub fn new(dim: usize, metric: Metric, bf16: bool) -> Result<Self> {
        let ctx = context()?;
        let mut h: *mut sys::trr_dense = std::ptr::null_mut();
        let dtype = if bf16 { sys::TRR_DTYPE_BF16 } else { sys::TRR_DTYPE_F32 };
        // SAFETY: `ctx` is a live context, `h` a valid out pointer
        check(unsafe { sys::trr_dense_create(ctx, dim as u32, metric.raw(), dtype, 0, &mut h) })?;
        Ok(Self { h, dim, id_of: Vec::new(), ord_of: HashMap::new() })
    }

    pub fn dimension(&self) -> usize {
        self.dim
    }

    /// `VectorStore::len` (src/index.rs:428-430)
    pub fn len(&self) -> usize {
        self.ord_of.len()
    }

    pub fn is_empty(&self) -> bool {
        self.ord_of.is_empty()
    }

    /// `VectorStore::insert` (src/index.rs:359-375).  An id that is already stored is replaced, like `HashMap::insert`.
    pub fn insert(&mut self, id: u128, embedding: &[f32]) -> Result<()> {
        self.insert_batch(&[id], embedding)
    }

    /// `VectorStore::insert_batch` (src/index.rs:378-383): `rows` holds `ids.len()` embeddings back to back.
    pub fn insert_batch(&mut self, ids: &[u128], rows: &[f32]) -> Result<()> {
        if ids.is_empty() {
            return Ok(());
        }
        if rows.len() != ids.len() * self.dim {
            return Err(DeviceError::DimensionMismatch { expected: self.dim, actual: rows.len() / ids.len() });
        }
        for id in ids {
            if let Some(&old) = self.ord_of.get(id) {
                // SAFETY: `h` is live; `old` is an ordinal this store handed out
                check(unsafe { sys::trr_dense_remove(self.h, old) })?;
            }
        }
        // SAFETY: `rows` holds ids.len() * dim floats
        check(unsafe { sys::trr_dense_append(self.h, rows.as_ptr(), ids.len() as u64) })?;
        for id in ids {
            let ord = self.id_of.len() as u32;
            self.id_of.push(*id);
            self.ord_of.insert(*id, ord);
        }
        Ok(())
    }

    /// `VectorStore::remove` (src/index.rs:421-424): true if the id was stored
    pub fn remove(&mut self, id: u128) -> Result<bool> {
        match self.ord_of.remove(&id) {
            None => Ok(false),
            Some(ord) => {
                // SAFETY: as in insert_batch
                check(unsafe { sys::trr_dense_remove(self.h, ord) })?;
                Ok(true)
            }
        }
    }

    /// `VectorStore::search` (src/index.rs:386-412)
    pub fn search(&self, query: &[f32], k: usize) -> Result<Vec<(u128, f32)>> {
        Ok(self.search_batch(query, 1, k)?.pop().unwrap_or_default())
    }

    /// Additive: `b` queries in one call (the reference has no batch API; batches are where the tensor cores pay).
    pub fn search_batch(&self, queries: &[f32], b: usize, k: usize) -> Result<Vec<Vec<(u128, f32)>>> {
        if b == 0 {
            return Ok(Vec::new());
        }
        if queries.len() != b * self.dim {
            return Err(DeviceError::DimensionMismatch { expected: self.dim, actual: queries.len() / b }); // :387-392
        }
        let k = k.min(self.id_of.len());
        if k == 0 {
            return Ok(vec![Vec::new(); b]);
        }
        let mut ord = vec![0u32; b * k];
        let mut score = vec![0f32; b * k];
        let mut n = vec![0u32; b];
        // SAFETY: the out buffers hold b * k (ord, score) and b (n) elements
        check(unsafe {
            sys::trr_dense_search(self.h, queries.as_ptr(), b as u32, k as u32, ord.as_mut_ptr(), score.as_mut_ptr(), n.as_mut_ptr())
        })?;
        Ok((0..b)
            .map(|q| (0..n[q] as usize).map(|i| (self.id_of[ord[q * k + i] as usize], score[q * k + i])).collect())
            .collect())
    }

    /// ordinal -> id (used by `hybrid_search`)
    pub fn id_of(&self, ordinal: u32) -> u128 {
        self.id_of[ordinal as usize]
    }

    /// Device snapshot (the reference's `VectorStore` is not serialisable; additive)
    pub fn save(&self, path: &str) -> Result<()> {
        let c = std::ffi::CString::new(path).map_err(|e| DeviceError::InvalidArg(e.to_string()))?;
        // SAFETY: `c` is NUL-terminated
        check(unsafe { sys::trr_dense_save(self.h, c.as_ptr()) })
    }
}

impl Drop for DenseStore {
    fn drop(&mut self) {
        // SAFETY: `h` came from trr_dense_create and is destroyed once
        unsafe { sys::trr_dense_destroy(self.h) };
    }
}

// ------------------------------------------------------------------------------------------------
// BM25Index
// ------------------------------------------------------------------------------------------------

/// The frozen, device-resident form of a `BM25Index`: CSR postings with impacts computed in the reference's operation
/// order (src/index.rs:137-153), plus the term dictionary.  `BM25Index` keeps it in a `#[serde(skip)] OnceCell` and drops
/// it on `add` / `remove` (or calls [`Bm25Device::append`] / [`Bm25Device::remove`] to update it in place).
pub struct Bm25Device {
    h: *mut sys::trr_bm25,
    term_id: HashMap<String, u32>,
    id_of: Vec<u128>,
}

// SAFETY: as for DenseStore
unsafe impl Send for Bm25Device {}
unsafe impl Sync for Bm25Device {}

/// `idf` exactly as the reference computes it (src/index.rs:147): f32 arithmetic, `f32::ln`
pub fn idf(doc_count: u32, df: u32) -> f32 {
    let (n, df) = (doc_count as f32, df as f32);
    ((n - df + 0.5) / (df + 0.5) + 1.0).ln()
}

impl Bm25Device {
    /// Builds the device index from the reference's own fields (src/index.rs:32-51).
    ///
    /// `order` lists the chunk ids in insertion order (the shim keeps it next to `doc_lengths`; a deserialised index that
    /// lacks it passes the ids sorted) - it defines the ordinals and with them the tie order.
    #[allow(clippy::too_many_arguments)]
    pub fn freeze(
        inverted_index: &HashMap<String, Vec<(u128, u32)>>,
        doc_freqs: &HashMap<String, u32>,
        doc_lengths: &HashMap<u128, u32>,
        order: &[u128],
        doc_count: u32,
        avg_doc_length: f32,
        k1: f32,
        b: f32,
    ) -> Result<Self> {
        let ctx = context()?;
        let ord_of: HashMap<u128, u32> = order.iter().enumerate().map(|(i, id)| (*id, i as u32)).collect();
        let doc_len: Vec<u32> = order.iter().map(|id| doc_lengths.get(id).copied().unwrap_or(0)).collect(); // :144
        // deterministic term ids: ascending term string
        let mut terms: Vec<&String> = inverted_index.keys().collect();
        terms.sort();
        let mut term_id = HashMap::with_capacity(terms.len());
        let mut term_off = Vec::with_capacity(terms.len() + 1);
        let (mut post_doc, mut post_tf, mut idfs) = (Vec::new(), Vec::new(), Vec::with_capacity(terms.len()));
        term_off.push(0u64);
        for (t, term) in terms.iter().enumerate() {
            term_id.insert((*term).clone(), t as u32);
            // postings sorted by ordinal; the first posting of a chunk wins (term_frequency uses `find`, :128-132)
            let mut pl: Vec<(u32, u32)> =
                inverted_index[*term].iter().filter_map(|(id, tf)| ord_of.get(id).map(|o| (*o, *tf))).collect();
            pl.sort_by_key(|p| p.0); // stable
            pl.dedup_by_key(|p| p.0);
            for (o, tf) in pl {
                post_doc.push(o);
                post_tf.push(tf);
            }
            term_off.push(post_doc.len() as u64);
            idfs.push(idf(doc_count, doc_freqs.get(*term).copied().unwrap_or(0))); // :140
        }
        let mut h: *mut sys::trr_bm25 = std::ptr::null_mut();
        // SAFETY: the arrays have the lengths the header documents (term_off: n_terms + 1, postings: term_off[n_terms],
        // doc_len: n_docs, idf: n_terms)
        check(unsafe {
            sys::trr_bm25_build(
                ctx,
                order.len() as u32,
                terms.len() as u32,
                term_off.as_ptr(),
                post_doc.as_ptr(),
                post_tf.as_ptr(),
                doc_len.as_ptr(),
                avg_doc_length,
                k1,
                b,
                idfs.as_ptr(),
                0,
                &mut h,
            )
        })?;
        Ok(Self { h, term_id, id_of: order.to_vec() })
    }

    /// term -> id for a tokenised query; unknown terms map to `u32::MAX` (they score 0.0 and select no document)
    pub fn term_ids<S: AsRef<str>>(&self, tokens: &[S]) -> Vec<u32> {
        tokens.iter().map(|t| self.term_id.get(t.as_ref()).copied().unwrap_or(u32::MAX)).collect()
    }

    /// `BM25Index::search` (src/index.rs:212-243) for an already tokenised query (the tokenizer stays in Rust, :111-124)
    pub fn search(&self, term_ids: &[u32], k: usize) -> Result<Vec<(u128, f32)>> {
        let k = k.min(self.id_of.len());
        if term_ids.is_empty() || k == 0 {
            return Ok(Vec::new()); // :214-216
        }
        let off = [0u32, term_ids.len() as u32];
        let mut ord = vec![0u32; k];
        let mut score = vec![0f32; k];
        let mut n = 0u32;
        // SAFETY: one query; out buffers hold k elements
        check(unsafe {
            sys::trr_bm25_search(self.h, term_ids.as_ptr(), off.as_ptr(), 1, k as u32, ord.as_mut_ptr(), score.as_mut_ptr(), &mut n)
        })?;
        Ok((0..n as usize).map(|i| (self.id_of[ord[i] as usize], score[i])).collect())
    }
}

impl Drop for Bm25Device {
    fn drop(&mut self) {
        // SAFETY: `h` came from trr_bm25_build and is destroyed once
        unsafe { sys::trr_bm25_destroy(self.h) };
    }
}

// ------------------------------------------------------------------------------------------------
// FusionStrategy::fuse and HybridRetriever::retrieve
// ------------------------------------------------------------------------------------------------

/// `FusionStrategy::fuse` (src/fusion.rs:42-63) on the device: ids are numbered by first appearance (dense list, then
/// sparse list), which is also the tie order of the result.
pub fn fuse(strategy: Fusion, dense: &[(u128, f32)], sparse: &[(u128, f32)]) -> Result<Vec<(u128, f32)>> {
    let ctx = context()?;
    let mut num: HashMap<u128, u32> = HashMap::new();
    let mut ids: Vec<u128> = Vec::new();
    let mut number = |id: u128| -> u32 {
        *num.entry(id).or_insert_with(|| {
            ids.push(id);
            (ids.len() - 1) as u32
        })
    };
    let c = dense.len().max(sparse.len()).max(1);
    let (mut d_ord, mut d_sc, mut s_ord, mut s_sc) = (vec![0u32; c], vec![0f32; c], vec![0u32; c], vec![0f32; c]);
    for (i, (id, s)) in dense.iter().enumerate() {
        d_ord[i] = number(*id);
        d_sc[i] = *s;
    }
    for (i, (id, s)) in sparse.iter().enumerate() {
        s_ord[i] = number(*id);
        s_sc[i] = *s;
    }
    let (d_n, s_n) = ([dense.len() as u32], [sparse.len() as u32]);
    let k_out = (dense.len() + sparse.len()).max(1);
    let mut out_ord = vec![0u32; k_out];
    let mut out_fused = vec![0f32; k_out];
    let mut out_n = 0u32;
    let (kind, param) = strategy.raw();
    // SAFETY: list buffers hold C elements, out buffers k_out; the per-source score outputs are optional (NULL)
    check(unsafe {
        sys::trr_fuse(
            ctx,
            kind,
            param,
            d_ord.as_ptr(),
            d_sc.as_ptr(),
            d_n.as_ptr(),
            s_ord.as_ptr(),
            s_sc.as_ptr(),
            s_n.as_ptr(),
            1,
            c as u32,
            k_out as u32,
            out_ord.as_mut_ptr(),
            out_fused.as_mut_ptr(),
            std::ptr::null_mut(),
            std::ptr::null_mut(),
            &mut out_n,
        )
    })?;
    Ok((0..out_n as usize).map(|i| (ids[out_ord[i] as usize], out_fused[i])).collect())
}

/// One row of `HybridRetriever::retrieve`'s output: the fields of `RetrievalResult` the device knows
/// (src/retrieve.rs:13-24; the chunk itself is looked up by id in the store's `chunks` map, :204-213)
#[derive(Debug, Clone, PartialEq)]
pub struct HybridHit {
    pub id: u128,
    pub fused_score: f32,
    pub dense_score: Option<f32>,
    pub sparse_score: Option<f32>,
}

/// `HybridRetriever::retrieve` (src/retrieve.rs:175-220) in one library call: dense top-C, sparse top-C, fusion, take(k).
/// The dense store and the BM25 index must hold the same chunks in the same insertion order (`HybridRetriever::index`
/// adds to both, :156-164).
pub fn hybrid_search(
    dense: &DenseStore,
    bm25: &Bm25Device,
    query_embedding: &[f32],
    query_term_ids: &[u32],
    candidates_per_source: usize,
    strategy: Fusion,
    k: usize,
) -> Result<Vec<HybridHit>> {
    if query_embedding.len() != dense.dim {
        return Err(DeviceError::DimensionMismatch { expected: dense.dim, actual: query_embedding.len() });
    }
    let k = k.min(dense.id_of.len());
    if k == 0 {
        return Ok(Vec::new());
    }
    let off = [0u32, query_term_ids.len() as u32];
    let mut ord = vec![0u32; k];
    let (mut fused, mut ds, mut ss) = (vec![0f32; k], vec![0f32; k], vec![0f32; k]);
    let mut n = 0u32;
    let (kind, param) = strategy.raw();
    // SAFETY: one query; every out buffer holds k elements
    check(unsafe {
        sys::trr_hybrid_search(
            dense.h,
            bm25.h,
            query_embedding.as_ptr(),
            query_term_ids.as_ptr(),
            off.as_ptr(),
            1,
            candidates_per_source as u32,
            kind,
            param,
            k as u32,
            1,
            1,
            ord.as_mut_ptr(),
            fused.as_mut_ptr(),
            ds.as_mut_ptr(),
            ss.as_mut_ptr(),
            &mut n,
        )
    })?;
    let opt = |v: f32| if v.is_nan() { None } else { Some(v) };
    Ok((0..n as usize)
        .map(|i| HybridHit { id: dense.id_of(ord[i]), fused_score: fused[i], dense_score: opt(ds[i]), sparse_score: opt(ss[i]) })
        .collect())
}
