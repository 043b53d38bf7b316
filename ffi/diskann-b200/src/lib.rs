//! Safe wrapper over `diskann-b200-sys` (see INTEGRATION.md for how it plugs into the reference:
//! `GpuIndex` is the device snapshot a `layers::GpuFull<T>` mirrors into, `search_batch` is what a
//! `benchmark_core::search::Search` implementation (`GpuKNN`) calls once per query batch).
//! This image has no Rust toolchain: the crate is kept in step with the header by tools/gen_ffi.py
//! (the -sys half) and mirrors diskann_b200/index.py, which the tests drive through the same ABI.
use diskann_b200_sys as sys;
use std::os::raw::c_void;
use std::ptr;

/// `diskann_vector::distance::Metric` values (`#[repr(C)]`, metric.rs:8-20) — pass `metric as i32`.
#[repr(i32)]
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum Metric {
    Cosine = 0,
    InnerProduct = 1,
    L2 = 2,
    CosineNormalized = 3,
}

/// Element types with a device path (VectorRepr implementors the reference instantiates).
pub trait Element: Copy {
    const DTYPE: i32;
}
impl Element for f32 {
    const DTYPE: i32 = 0;
}
impl Element for i8 {
    const DTYPE: i32 = 2;
}
impl Element for u8 {
    const DTYPE: i32 = 3;
}
// f16: `half::f16` with DTYPE = 1 (the crate does not depend on `half`; add the impl next to it)

#[derive(Debug)]
pub struct Error(pub String);
pub type Result<T> = std::result::Result<T, Error>;

fn check(status: i32) -> Result<()> {
    sys::check(status).map_err(Error)
}

/// Results of one batched search: row-major `[nq][k]`, padded with `u32::MAX` / `+inf`;
/// `cmps` / `hops` follow `SearchStats` (diskann/src/graph/index.rs:1990-1991).
pub struct Batch {
    pub k: usize,
    pub ids: Vec<u32>,
    pub dists: Vec<f32>,
    pub counts: Vec<u32>,
    pub cmps: Vec<u32>,
    pub hops: Vec<u32>,
}

pub struct GpuIndex<T: Element> {
    raw: *mut sys::dab_index,
    dim: usize,
    _marker: std::marker::PhantomData<T>,
}

// one CUDA stream per handle; read-only calls from one thread at a time
unsafe impl<T: Element> Send for GpuIndex<T> {}

impl<T: Element> GpuIndex<T> {
    pub fn new(metric: Metric, dim: usize, n_points: u64, n_start: u32, max_degree: u32, device: i32) -> Result<Self> {
        let mut raw = ptr::null_mut();
        check(unsafe { sys::dab_create(&mut raw, T::DTYPE, metric as i32, dim as u32, n_points, n_start, max_degree, device) })?;
        Ok(Self { raw, dim, _marker: std::marker::PhantomData })
    }

    /// Dense row-major rows `[count][dim]` starting at row `first` (start points follow the data points).
    pub fn upload_vectors(&mut self, rows: &[T], first: u64) -> Result<()> {
        assert_eq!(rows.len() % self.dim, 0, "rows must be a whole number of vectors");
        check(unsafe { sys::dab_upload_vectors(self.raw, rows.as_ptr() as *const c_void, first, (rows.len() / self.dim) as u64) })
    }

    /// Adjacency rows `[count][stride]` with `row[0] = degree` (diskann-inmem/src/neighbors.rs:69-163).
    pub fn upload_graph(&mut self, adj: &[u32], stride: u32, first: u64) -> Result<()> {
        assert_eq!(adj.len() % stride as usize, 0);
        check(unsafe { sys::dab_upload_graph(self.raw, adj.as_ptr(), stride, first, (adj.len() / stride as usize) as u64) })
    }

    /// Batched `multi_insert`-style construction on the device over the uploaded vectors.
    pub fn build(&mut self, pruned_degree: u32, l_build: u32, alpha: f32) -> Result<()> {
        check(unsafe { sys::dab_build(self.raw, pruned_degree, l_build, alpha, 0) })
    }

    /// `KNN::search` for every query of the batch at once (search_internal + post-processing).
    pub fn search_batch(&self, queries: &[T], k: usize, l_search: u32, beam_width: u32) -> Result<Batch> {
        assert_eq!(queries.len() % self.dim, 0);
        let nq = queries.len() / self.dim;
        let mut b = Batch { k, ids: vec![0; nq * k], dists: vec![0.0; nq * k], counts: vec![0; nq], cmps: vec![0; nq], hops: vec![0; nq] };
        check(unsafe {
            sys::dab_search_batch(self.raw, queries.as_ptr() as *const c_void, nq as u32, k as u32, l_search, beam_width,
                                  b.ids.as_mut_ptr(), b.dists.as_mut_ptr(), b.counts.as_mut_ptr(), b.cmps.as_mut_ptr(), b.hops.as_mut_ptr())
        })?;
        Ok(b)
    }

    /// Queue a batch on `slot` without waiting (`search_all`'s one task per partition, api.rs:410-419, mapped to
    /// device slots).  The returned guard borrows the queries and owns the result buffers; `InFlight::wait` joins it.
    pub fn search_batch_async<'a>(&'a self, slot: u32, queries: &'a [T], k: usize, l_search: u32, beam_width: u32) -> Result<InFlight<'a, T>> {
        assert_eq!(queries.len() % self.dim, 0);
        let nq = queries.len() / self.dim;
        let mut b = Batch { k, ids: vec![0; nq * k], dists: vec![0.0; nq * k], counts: vec![0; nq], cmps: vec![0; nq], hops: vec![0; nq] };
        check(unsafe {
            sys::dab_search_batch_async(self.raw, slot, queries.as_ptr() as *const c_void, nq as u32, k as u32, l_search, beam_width,
                                        b.ids.as_mut_ptr(), b.dists.as_mut_ptr(), b.counts.as_mut_ptr(), b.cmps.as_mut_ptr(), b.hops.as_mut_ptr())
        })?;
        Ok(InFlight { index: self, slot, batch: Some(b), _queries: queries })
    }

    /// PQ traversal + the providers' full-precision `Rerank` (what `use_fp_for_search: false` runs).
    pub fn search_batch_pq_rerank(&self, queries: &[T], k: usize, l_search: u32, beam_width: u32) -> Result<Batch> {
        assert_eq!(queries.len() % self.dim, 0);
        let nq = queries.len() / self.dim;
        let mut b = Batch { k, ids: vec![0; nq * k], dists: vec![0.0; nq * k], counts: vec![0; nq], cmps: vec![0; nq], hops: vec![0; nq] };
        check(unsafe {
            sys::dab_search_batch_pq_rerank(self.raw, queries.as_ptr() as *const c_void, nq as u32, k as u32, l_search, beam_width,
                                            b.ids.as_mut_ptr(), b.dists.as_mut_ptr(), b.counts.as_mut_ptr(), b.cmps.as_mut_ptr(),
                                            b.hops.as_mut_ptr())
        })?;
        Ok(b)
    }

    /// `train_pq` + encoding of every stored row, on the device.
    pub fn train_pq(&mut self, train: &[f32], n_chunks: u32, seed: u64) -> Result<()> {
        assert_eq!(train.len() % self.dim, 0);
        check(unsafe { sys::dab_pq_train(self.raw, train.as_ptr(), (train.len() / self.dim) as u64, n_chunks, 256, 5, seed) })?;
        check(unsafe { sys::dab_pq_encode_all(self.raw) })
    }

    /// The scalar-quantized store (`SQStore<NBITS>`): hand over the quantizer, encode every resident row on the device.
    pub fn set_scalar_quantizer(&mut self, nbits: i32, shift: &[f32], scale: f32, shift_square_norm: f32, mean_norm: Option<f32>) -> Result<()> {
        assert_eq!(shift.len(), self.dim);
        check(unsafe {
            sys::dab_upload_sq(self.raw, nbits, shift.as_ptr(), scale, shift_square_norm, mean_norm.unwrap_or(0.0), std::ptr::null())
        })?;
        check(unsafe { sys::dab_sq_encode_all(self.raw) })
    }

    /// Traversal over the scalar-quantized rows; `rerank` adds `Pipeline<FilterStartPoints, Rerank>`.
    pub fn search_batch_sq(&self, queries: &[T], k: usize, l_search: u32, beam_width: u32, rerank: bool) -> Result<Batch> {
        assert_eq!(queries.len() % self.dim, 0);
        let nq = queries.len() / self.dim;
        let mut b = Batch { k, ids: vec![0; nq * k], dists: vec![0.0; nq * k], counts: vec![0; nq], cmps: vec![0; nq], hops: vec![0; nq] };
        check(unsafe {
            sys::dab_search_batch_sq(self.raw, queries.as_ptr() as *const c_void, nq as u32, k as u32, l_search, beam_width, rerank as i32,
                                     b.ids.as_mut_ptr(), b.dists.as_mut_ptr(), b.counts.as_mut_ptr(), b.cmps.as_mut_ptr(),
                                     b.hops.as_mut_ptr())
        })?;
        Ok(b)
    }

    /// One process per GPU: join the communicator described by `id` (from `unique_id()` on rank 0) …
    pub fn comm_init(&mut self, id: &[u8; 128], n_ranks: i32, rank: i32) -> Result<()> {
        check(unsafe { sys::dab_comm_init(self.raw, id.as_ptr() as *const _, n_ranks, rank) })
    }

    /// … and replicate the resident snapshot from `root` (one NCCL broadcast per buffer, at load).
    pub fn broadcast_index(&mut self, root: i32) -> Result<()> {
        check(unsafe { sys::dab_broadcast_index(self.raw, root) })
    }
}

/// A batch in flight on one slot of the device.  Dropping it joins the slot (the library writes into the
/// buffers it owns until then).
pub struct InFlight<'a, T: Element> {
    index: &'a GpuIndex<T>,
    slot: u32,
    batch: Option<Batch>,
    _queries: &'a [T],
}

impl<'a, T: Element> InFlight<'a, T> {
    pub fn wait(mut self) -> Result<Batch> {
        check(unsafe { sys::dab_wait(self.index.raw, self.slot) })?;
        Ok(self.batch.take().expect("joined once"))
    }
}

impl<'a, T: Element> Drop for InFlight<'a, T> {
    fn drop(&mut self) {
        if self.batch.is_some() {
            unsafe { sys::dab_wait(self.index.raw, self.slot) };
        }
    }
}

pub fn unique_id() -> Result<[u8; 128]> {
    let mut id = [0u8; 128];
    check(unsafe { sys::dab_comm_unique_id(id.as_mut_ptr() as *mut _) })?;
    Ok(id)
}

impl<T: Element> Drop for GpuIndex<T> {
    fn drop(&mut self) {
        unsafe { sys::dab_destroy(self.raw) }
    }
}

/// `MinMaxQuantizer::new(Transform::Null(dim), grid_scale)` + `compress_into` for `vectors.len() / dim` vectors:
/// rows in the canonical-front layout of `minmax::Data<NBITS>` (`DataRef::from_canonical_front(&row, dim)`), and the
/// `L2Loss` of every vector.  `Err` when an input vector contains NaN (`InputContainsNaN`).
pub fn minmax_compress(device: i32, grid_scale: f32, dim: usize, nbits: i32, vectors: &[f32]) -> Result<(Vec<u8>, Vec<f32>)> {
    let n = vectors.len() / dim;
    let row_bytes = unsafe { sys::dab_minmax_row_bytes(dim as u32, nbits) } as usize;
    let mut rows = vec![0u8; n * row_bytes];
    let mut loss = vec![0f32; n];
    check(unsafe {
        sys::dab_minmax_compress(device, grid_scale, dim as u32, nbits, vectors.as_ptr(), n as u64, rows.as_mut_ptr(), loss.as_mut_ptr())
    })?;
    Ok((rows, loss))
}

/// `MinMax{L2Squared, IP, Cosine, CosineNormalized}::evaluate(DataRef<N>, DataRef<M>)` for row pairs (N x N or 8 x N bits).
pub fn minmax_distances(device: i32, metric: Metric, nbits_x: i32, nbits_y: i32, dim: usize, x_rows: &[u8], y_rows: &[u8]) -> Result<Vec<f32>> {
    let n = x_rows.len() / unsafe { sys::dab_minmax_row_bytes(dim as u32, nbits_x) } as usize;
    let mut out = vec![0f32; n];
    check(unsafe {
        sys::dab_minmax_distances(device, metric as i32, nbits_x, nbits_y, dim as u32, x_rows.as_ptr(), y_rows.as_ptr(), n as u64, out.as_mut_ptr())
    })?;
    Ok(out)
}

/// Full-precision queries (`FullQuery`) against compressed rows: `MinMax*::evaluate(FullQueryRef, DataRef<NBITS>)` for
/// every (query, row) pair, row-major `[nq][n]`.
pub fn minmax_query_distances(device: i32, metric: Metric, nbits: i32, dim: usize, queries: &[f32], rows: &[u8]) -> Result<Vec<f32>> {
    let nq = queries.len() / dim;
    let n = rows.len() / unsafe { sys::dab_minmax_row_bytes(dim as u32, nbits) } as usize;
    let mut out = vec![0f32; nq * n];
    check(unsafe {
        sys::dab_minmax_query_distances(device, metric as i32, nbits, dim as u32, queries.as_ptr(), nq as u32, rows.as_ptr(), n as u64, out.as_mut_ptr())
    })?;
    Ok(out)
}
