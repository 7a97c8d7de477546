//! Rust binding of include/bsr.h -- SOURCE ONLY (no Rust toolchain in the build image).
//!
//! `ffi` is the raw `extern "C"` block; `GpuVectorstore` mirrors the reference's
//! `PolarsVectorstore` (src/vectorstore/polars.rs) with the rows resident in one GPU's HBM, and
//! `compute_local_top_k` / `compute_global_top_k` mirror src/mpi_helpers/metrics.rs.

use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};

pub mod ffi {
    use super::*;

    #[repr(C)]
    pub struct bsr_store {
        _private: [u8; 0],
    }

    #[repr(C)]
    pub struct bsr_comm {
        _private: [u8; 0],
    }

    pub const BSR_COMM_ID_BYTES: usize = 128;
    pub const BSR_OK: c_int = 0;
    pub const BSR_F32: c_int = 0;
    pub const BSR_BF16: c_int = 1;
    pub const BSR_COSINE: c_int = 0;
    pub const BSR_DOT: c_int = 1;
    pub const BSR_MAX_TOP_K: u32 = 1024;

    extern "C" {
        pub fn bsr_version() -> c_int;
        pub fn bsr_last_error() -> *const c_char;
        pub fn bsr_device_count(n_devices: *mut c_int) -> c_int;

        pub fn bsr_store_create(device: c_int, dim: u32, dtype: c_int, row_base: u64,
                                capacity_rows: u64, out: *mut *mut bsr_store) -> c_int;
        pub fn bsr_store_destroy(store: *mut bsr_store) -> c_int;
        pub fn bsr_store_reset(store: *mut bsr_store) -> c_int;
        pub fn bsr_store_append(store: *mut bsr_store, rows: *const f32, n_rows: u64) -> c_int;
        pub fn bsr_store_append_device(store: *mut bsr_store, d_rows: *const f32, n_rows: u64) -> c_int;
        pub fn bsr_store_append_synthetic(store: *mut bsr_store, seed: u64, first_row: u64,
                                          n_rows: u64) -> c_int;
        pub fn bsr_store_set_metric(store: *mut bsr_store, metric: c_int) -> c_int;
        pub fn bsr_store_metric(store: *const bsr_store, metric: *mut c_int) -> c_int;
        pub fn bsr_store_count(store: *const bsr_store, n_rows: *mut u64) -> c_int;
        pub fn bsr_store_dim(store: *const bsr_store, dim: *mut u32) -> c_int;
        pub fn bsr_store_row_base(store: *const bsr_store, row_base: *mut u64) -> c_int;
        pub fn bsr_store_device(store: *const bsr_store, device: *mut c_int) -> c_int;
        pub fn bsr_store_scan_bytes(store: *const bsr_store, bytes: *mut u64) -> c_int;
        pub fn bsr_store_get_many(store: *const bsr_store, offset: u64, length: u64,
                                  out_rows: *mut f32) -> c_int;

        pub fn bsr_search(store: *mut bsr_store, queries: *const f32, n_queries: u32, top_k: u32,
                          out_idx: *mut u64, out_dist: *mut f32, out_count: *mut u32) -> c_int;
        pub fn bsr_candidate_slots(top_k: u32) -> u32;
        pub fn bsr_search_device(store: *mut bsr_store, d_queries: *const f32, n_queries: u32,
                                 top_k: u32, d_candidates: *mut u64, stream: *mut c_void) -> c_int;
        pub fn bsr_merge_device(store: *mut bsr_store, d_lists: *const u64, n_lists: u32,
                                n_queries: u32, top_k: u32, d_out: *mut u64,
                                stream: *mut c_void) -> c_int;
        pub fn bsr_unpack_device(store: *mut bsr_store, d_candidates: *const u64, n_queries: u32,
                                 top_k: u32, d_idx: *mut u64, d_dist: *mut f32, d_count: *mut u32,
                                 stream: *mut c_void) -> c_int;

        pub fn bsr_merge_host(store: *mut bsr_store, idx: *const u64, dist: *const f32, n: u64,
                              top_k: u32, out_idx: *mut u64, out_dist: *mut f32,
                              out_count: *mut u32) -> c_int;
        pub fn bsr_comm_unique_id(id: *mut u8) -> c_int;
        pub fn bsr_comm_create(id: *const u8, n_ranks: c_int, rank: c_int, device: c_int,
                               out: *mut *mut bsr_comm) -> c_int;
        pub fn bsr_comm_destroy(comm: *mut bsr_comm) -> c_int;
        pub fn bsr_search_parallel(shard: *mut bsr_store, comm: *mut bsr_comm, queries: *const f32,
                                   n_queries: u32, top_k: u32, out_idx: *mut u64, out_dist: *mut f32,
                                   out_count: *mut u32) -> c_int;
        pub fn bsr_search_parallel_device(shard: *mut bsr_store, comm: *mut bsr_comm, d_queries: *const f32,
                                          n_queries: u32, top_k: u32, d_candidates: *mut u64,
                                          stream: *mut c_void) -> c_int;
        pub fn bsr_comm_uses_mailbox(comm: *const bsr_comm, yes: *mut c_int) -> c_int;
        pub fn bsr_store_set_tensor_min_batch(store: *mut bsr_store, min_batch: u32) -> c_int;
        pub fn bsr_store_set_option(store: *mut bsr_store, name: *const c_char, value: i64) -> c_int;
        pub fn bsr_store_launch_count(store: *const bsr_store, n_launches: *mut u64) -> c_int;
        pub fn bsr_store_last_scan_ms(store: *mut bsr_store, ms: *mut f32) -> c_int;
        pub fn bsr_store_set_profiling(store: *mut bsr_store, enabled: c_int) -> c_int;
        pub fn bsr_store_scan_profile(store: *mut bsr_store, total_ms: *mut f64,
                                      n_scans: *mut u64) -> c_int;
    }
}

/// Error carrying the status code and the library's thread-local message.
#[derive(Debug, Clone)]
pub struct BsrError {
    pub status: i32,
    pub message: String,
}

impl std::fmt::Display for BsrError {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        write!(f, "bsr status {}: {}", self.status, self.message)
    }
}
impl std::error::Error for BsrError {}

fn check(status: c_int) -> Result<(), BsrError> {
    if status == ffi::BSR_OK {
        return Ok(());
    }
    let message = unsafe {
        let p = ffi::bsr_last_error();
        if p.is_null() { String::new() } else { CStr::from_ptr(p).to_string_lossy().into_owned() }
    };
    Err(BsrError { status, message })
}

/// src/vectorstore/polars.rs:12-15
pub struct SliceArgs {
    pub offset: i32,
    pub length: usize,
}

/// `PolarsVectorstore` (src/vectorstore/polars.rs:7-10) with the rows in one GPU's HBM.
/// Parquet decode/encode stays with polars in the caller (`new`/`reload`/`persist` there would
/// call `append_many` / `get_many` below).
pub struct GpuVectorstore {
    handle: *mut ffi::bsr_store,
    dim: u32,
}

// one handle is used from one host thread at a time (include/bsr.h)
unsafe impl Send for GpuVectorstore {}

impl GpuVectorstore {
    /// `PolarsVectorstore::new(path, empty = true)`; `row_base` is the rank's
    /// `RankInterval::start_index` (src/mpi_helpers/load_balance.rs:24-42).
    pub fn new_empty(device: i32, dim: u32, row_base: u64, capacity_rows: u64) -> Result<Self, BsrError> {
        let mut handle = std::ptr::null_mut();
        check(unsafe {
            ffi::bsr_store_create(device, dim, ffi::BSR_F32, row_base, capacity_rows, &mut handle)
        })?;
        Ok(Self { handle, dim })
    }

    /// polars.rs:93-95
    /// `dot = false`: the reference's cosine_distance (default); `dot = true`: 1 - dot product
    /// (BSR_DOT, for stores normalised at ingest).
    pub fn set_dot_metric(&mut self, dot: bool) -> Result<(), BsrError> {
        check(unsafe { ffi::bsr_store_set_metric(self.handle, if dot { ffi::BSR_DOT } else { ffi::BSR_COSINE }) })
    }

    /// Ablation switch / test hook of this handle (`bsr_store_set_option`, include/bsr.h).
    pub fn set_option(&mut self, name: &str, value: i64) -> Result<(), BsrError> {
        let c = std::ffi::CString::new(name).map_err(|_| BsrError { status: -1, message: "bad option name".into() })?;
        check(unsafe { ffi::bsr_store_set_option(self.handle, c.as_ptr(), value) })
    }

    pub fn reset(&mut self) -> Result<(), BsrError> {
        check(unsafe { ffi::bsr_store_reset(self.handle) })
    }

    /// polars.rs:97-99
    pub fn append(&mut self, vector: &Vec<f32>) -> Result<(), BsrError> {
        assert_eq!(vector.len(), self.dim as usize);
        check(unsafe { ffi::bsr_store_append(self.handle, vector.as_ptr(), 1) })
    }

    /// polars.rs:101-119
    pub fn append_many(&mut self, vectors: &Vec<Vec<f32>>) -> Result<(), BsrError> {
        if vectors.is_empty() {
            return Ok(());
        }
        let mut flat = Vec::with_capacity(vectors.len() * self.dim as usize);
        for v in vectors {
            assert_eq!(v.len(), self.dim as usize);
            flat.extend_from_slice(v);
        }
        check(unsafe { ffi::bsr_store_append(self.handle, flat.as_ptr(), vectors.len() as u64) })
    }

    /// polars.rs:243-246
    pub fn get_count(&self) -> Result<usize, BsrError> {
        let mut n = 0u64;
        check(unsafe { ffi::bsr_store_count(self.handle, &mut n) })?;
        Ok(n as usize)
    }

    /// polars.rs:121-156 (the slice is clamped to the store, as DataFrame::slice does)
    pub fn get_many(&self, slice: Option<SliceArgs>) -> Result<Vec<Vec<f32>>, BsrError> {
        let count = self.get_count()? as i64;
        let (start, len) = match slice {
            None => (0i64, count),
            Some(a) => {
                let off = a.offset as i64;
                let start = if off < 0 { (count + off).max(0) } else { off.min(count) };
                (start, (a.length as i64).min(count - start).max(0))
            }
        };
        let d = self.dim as usize;
        let mut flat = vec![0f32; len as usize * d];
        if len > 0 {
            check(unsafe {
                ffi::bsr_store_get_many(self.handle, start as u64, len as u64, flat.as_mut_ptr())
            })?;
        }
        Ok(flat.chunks(d).map(|c| c.to_vec()).collect())
    }

    /// polars.rs:158-169
    pub fn get(&self, index: usize) -> Result<Vec<f32>, BsrError> {
        self.get_many(Some(SliceArgs { offset: index as i32, length: 1 }))?
            .into_iter()
            .next()
            .ok_or(BsrError { status: -4, message: "Index not found".into() })
    }

    /// `compute_local_top_k` (src/mpi_helpers/metrics.rs:16-53) over this shard:
    /// (global index, distance) ascending, at most `top_k` long.
    pub fn compute_local_top_k(&mut self, top_k: usize, target_vector: &Vec<f32>)
        -> Result<Vec<(usize, f32)>, BsrError> {
        if self.get_count()? == 0 {
            return Ok(vec![]);
        }
        if target_vector.len() != self.dim as usize {
            // cosine_distance is 1.0 for every row when lengths differ (src/metrics.rs:144-146)
            let n = self.get_count()?.min(top_k);
            let mut base = 0u64;
            check(unsafe { ffi::bsr_store_row_base(self.handle, &mut base) })?;
            return Ok((0..n).map(|i| (base as usize + i, 1.0f32)).collect());
        }
        let mut idx = vec![0u64; top_k];
        let mut dist = vec![0f32; top_k];
        let mut count = 0u32;
        check(unsafe {
            ffi::bsr_search(self.handle, target_vector.as_ptr(), 1, top_k as u32,
                            idx.as_mut_ptr(), dist.as_mut_ptr(), &mut count)
        })?;
        Ok((0..count as usize).map(|i| (idx[i] as usize, dist[i])).collect())
    }
}

/// `compute_global_top_k` (src/mpi_helpers/metrics.rs:141-171): the gathered pairs are sorted by
/// (distance, index) on the GPU; `scratch` is any store handle on the device that should do it.
pub fn compute_global_top_k(scratch: &mut GpuVectorstore, global_indices: Vec<usize>,
                            global_distances: Vec<f32>, top_k: usize)
    -> Result<Vec<(usize, f32)>, BsrError> {
    if global_indices.is_empty() {
        return Ok(vec![]);
    }
    let gi: Vec<u64> = global_indices.iter().map(|&i| i as u64).collect();
    let mut idx = vec![0u64; top_k];
    let mut dist = vec![0f32; top_k];
    let mut count = 0u32;
    check(unsafe {
        ffi::bsr_merge_host(scratch.handle, gi.as_ptr(), global_distances.as_ptr(), gi.len() as u64,
                            top_k as u32, idx.as_mut_ptr(), dist.as_mut_ptr(), &mut count)
    })?;
    Ok((0..count as usize).map(|i| (idx[i] as usize, dist[i])).collect())
}

/// The library's own NCCL communicator (one per rank = per process = per GPU).
pub struct DeviceWorld {
    handle: *mut ffi::bsr_comm,
    pub rank: i32,
    pub size: i32,
}

impl DeviceWorld {
    /// Rank 0 creates the id; broadcast the 128 bytes with the host's own channel
    /// (`world.process_at_rank(ROOT).broadcast_into(&mut id)` in the reference's rsmpi world).
    pub fn unique_id() -> Result<[u8; ffi::BSR_COMM_ID_BYTES], BsrError> {
        let mut id = [0u8; ffi::BSR_COMM_ID_BYTES];
        check(unsafe { ffi::bsr_comm_unique_id(id.as_mut_ptr()) })?;
        Ok(id)
    }
    pub fn new(id: &[u8; ffi::BSR_COMM_ID_BYTES], size: i32, rank: i32, device: i32) -> Result<Self, BsrError> {
        let mut handle = std::ptr::null_mut();
        check(unsafe { ffi::bsr_comm_create(id.as_ptr(), size, rank, device, &mut handle) })?;
        Ok(Self { handle, rank, size })
    }
    /// True when single-query searches with top_k <= 128 exchange their candidates through NVLink peer
    /// stores fused into the scan kernel (one kernel per rank and query) instead of ncclAllGather.
    pub fn uses_mailbox(&self) -> Result<bool, BsrError> {
        let mut yes: c_int = 0;
        check(unsafe { ffi::bsr_comm_uses_mailbox(self.handle, &mut yes) })?;
        Ok(yes != 0)
    }
}

impl Drop for DeviceWorld {
    fn drop(&mut self) {
        unsafe { ffi::bsr_comm_destroy(self.handle) };
    }
}

/// `parallel_top_k_similarity_search` (src/mpi_helpers/metrics.rs:174-206): `Some` on root.
pub fn parallel_top_k_similarity_search(world: &DeviceWorld, shard: &mut GpuVectorstore, top_k: usize,
                                        target_vector: &Vec<f32>)
    -> Result<Option<Vec<(usize, f32)>>, BsrError> {
    let root = world.rank == 0;
    let mut idx = vec![0u64; top_k];
    let mut dist = vec![0f32; top_k];
    let mut count = 0u32;
    check(unsafe {
        ffi::bsr_search_parallel(shard.handle, world.handle, target_vector.as_ptr(), 1, top_k as u32,
                                 if root { idx.as_mut_ptr() } else { std::ptr::null_mut() },
                                 if root { dist.as_mut_ptr() } else { std::ptr::null_mut() },
                                 if root { &mut count } else { std::ptr::null_mut() })
    })?;
    if !root {
        return Ok(None);
    }
    Ok(Some((0..count as usize).map(|i| (idx[i] as usize, dist[i])).collect()))
}

impl Drop for GpuVectorstore {
    fn drop(&mut self) {
        unsafe { ffi::bsr_store_destroy(self.handle) };
    }
}

/// src/mpi_helpers/load_balance.rs:8-42, unchanged arithmetic.
pub struct RankInterval {
    pub start_index: usize,
    pub end_index: usize,
}

pub fn interval_by_rank(rank: i32, size: i32, count: usize) -> RankInterval {
    let per_rank = if size as usize > count { 1 } else { count.div_ceil(size as usize) };
    let start_index = per_rank * rank as usize;
    let end_index = if rank == size - 1 { count } else { std::cmp::min(start_index + per_rank, count) };
    RankInterval { start_index, end_index }
}
