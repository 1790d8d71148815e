//! Raw declarations of `include/bpp_b200.h` (keep in step with that header; `tests/test_abi_host.py` checks the header against
//! the symbols the shared object exports).  Conventions: caller-owned buffers, little-endian, scalars and points as 32-byte
//! encodings, every function returns a `bpp_status` (0..5 mirror `ProofError`, >= 100 are CUDA / runtime failures).
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_void};

pub const BPP_OK: i32 = 0;
pub const BPP_VERIFICATION_FAILED: i32 = 1;
pub const BPP_INVALID_ARGUMENT: i32 = 2;
pub const BPP_INVALID_LENGTH: i32 = 3;
pub const BPP_INVALID_BLAKE2B: i32 = 4;
pub const BPP_SIZE_OVERFLOW: i32 = 5;
pub const BPP_ERR_CUDA: i32 = 100;
pub const BPP_ERR_INTERNAL: i32 = 101;

pub const BPP_RECOVER_ONLY: i32 = 0;
pub const BPP_RECOVER_AND_VERIFY: i32 = 1;
pub const BPP_VERIFY_ONLY: i32 = 2;

pub const BPP_MAX_BATCH: usize = 256;
pub const BPP_TRANSCRIPT_BYTES: usize = 203;

#[repr(C)]
pub struct bpp_ctx {
    _private: [u8; 0],
}
#[repr(C)]
pub struct bpp_gens {
    _private: [u8; 0],
}
#[repr(C)]
pub struct bpp_vbatch {
    _private: [u8; 0],
}
#[repr(C)]
pub struct bpp_vqueue {
    _private: [u8; 0],
}

/// `bpp_verify_args` (K reference calls = "chunks" in one device pass)
#[repr(C)]
pub struct bpp_verify_args {
    pub n_proofs: usize,
    pub n_chunks: usize,
    pub chunk_offsets: *const u64,
    pub proof_bytes: *const u8,
    pub proof_offsets: *const u64,
    pub commitments32: *const u8,
    pub commit_offsets: *const u64,
    pub min_values: *const u64,
    pub min_present: *const u8,
    pub seed_nonces32: *const u8,
    pub seed_present: *const u8,
    pub transcripts: *mut u8,
    pub action: i32,
}

/// `bpp_verify_challenges` (challenge-input form: the caller keeps merlin)
#[repr(C)]
pub struct bpp_verify_challenges {
    pub challenges32: *const u8,
    pub challenge_offsets: *const u64,
    pub weights32: *const u8,
}

#[repr(C)]
pub struct bpp_prove_args {
    pub n_proofs: usize,
    pub aggregation: i32,
    pub commitments32: *const u8,
    pub values: *const u64,
    pub blindings32: *const u8,
    pub min_values: *const u64,
    pub min_present: *const u8,
    pub seed_nonces32: *const u8,
    pub seed_present: *const u8,
    pub transcripts: *mut u8,
    pub rng_bytes: *const u8,
    pub rng_stride: usize,
}

extern "C" {
    // context
    pub fn bpp_ctx_create(device_ordinal: i32, out: *mut *mut bpp_ctx) -> i32;
    pub fn bpp_ctx_destroy(ctx: *mut bpp_ctx);
    pub fn bpp_last_error(ctx: *const bpp_ctx) -> *const c_char;
    pub fn bpp_ctx_sync(ctx: *mut bpp_ctx) -> i32;
    pub fn bpp_ctx_launch_count(ctx: *const bpp_ctx) -> u64;
    pub fn bpp_ctx_stream(ctx: *mut bpp_ctx) -> *mut c_void;
    pub fn bpp_ctx_set_replay_mode(ctx: *mut bpp_ctx, on_device: i32) -> i32;
    pub fn bpp_ctx_set_graphs(ctx: *mut bpp_ctx, enable: i32) -> i32;
    pub fn bpp_ctx_set_throughput_mode(ctx: *mut bpp_ctx, enable: i32) -> i32;
    pub fn bpp_ctx_set_host_threads(ctx: *mut bpp_ctx, n: i32) -> i32;

    // batched point primitives
    pub fn bpp_decompress_check(ctx: *mut bpp_ctx, n: usize, in32: *const u8, ok: *mut u8, out32_or_null: *mut u8) -> i32;
    pub fn bpp_from_uniform_batch(ctx: *mut bpp_ctx, n: usize, in64: *const u8, out32: *mut u8) -> i32;
    pub fn bpp_points_sum_host(n: usize, in32: *const u8, out32: *mut u8) -> i32;

    // multiscalar multiplication
    pub fn bpp_msm(ctx: *mut bpp_ctx, n: usize, scalars32: *const u8, points32: *const u8, out32: *mut u8) -> i32;
    pub fn bpp_msm_segmented(ctx: *mut bpp_ctx, k: usize, offsets: *const u64, scalars32: *const u8, points32: *const u8, out32: *mut u8) -> i32;

    // generators
    pub fn bpp_gens_create(ctx: *mut bpp_ctx, bit_length: i32, max_aggregation: i32, extension_degree: i32, out: *mut *mut bpp_gens) -> i32;
    pub fn bpp_gens_create_with_bases(
        ctx: *mut bpp_ctx,
        bit_length: i32,
        max_aggregation: i32,
        extension_degree: i32,
        h_base32_or_null: *const u8,
        g_bases32_or_null: *const u8,
        out: *mut *mut bpp_gens,
    ) -> i32;
    pub fn bpp_gens_destroy(g: *mut bpp_gens);
    pub fn bpp_gens_get(g: *const bpp_gens, which: i32, index: usize, out32: *mut u8) -> i32;
    pub fn bpp_pedersen_commit_batch(g: *mut bpp_gens, count: usize, values: *const u64, blindings32: *const u8, n_blindings: i32, out32: *mut u8) -> i32;

    // batch verification
    pub fn bpp_verify_chunks(g: *mut bpp_gens, args: *const bpp_verify_args, chunk_status: *mut i32, masks32: *mut u8, mask_present: *mut u8) -> i32;
    pub fn bpp_verify_chunks_ch(
        g: *mut bpp_gens,
        args: *const bpp_verify_args,
        ch: *const bpp_verify_challenges,
        chunk_status: *mut i32,
        masks32: *mut u8,
        mask_present: *mut u8,
    ) -> i32;
    pub fn bpp_vbatch_create(g: *mut bpp_gens, args: *const bpp_verify_args, out: *mut *mut bpp_vbatch) -> i32;
    pub fn bpp_vbatch_create_multi(g: *mut bpp_gens, n_calls: usize, calls: *const *const bpp_verify_args, out: *mut *mut bpp_vbatch) -> i32;
    pub fn bpp_vbatch_run(vb: *mut bpp_vbatch, chunk_status: *mut i32, masks32: *mut u8, mask_present: *mut u8) -> i32;
    pub fn bpp_vbatch_transcripts(vb: *const bpp_vbatch, transcripts: *mut u8) -> i32;
    pub fn bpp_vbatch_destroy(vb: *mut bpp_vbatch);

    // coalescing queue
    pub fn bpp_vqueue_create(
        device_ordinal: i32,
        bit_length: i32,
        max_aggregation: i32,
        extension_degree: i32,
        h_base32_or_null: *const u8,
        g_bases32_or_null: *const u8,
        lanes: i32,
        max_calls_per_pass: i32,
        host_threads_per_lane: i32,
        out: *mut *mut bpp_vqueue,
    ) -> i32;
    pub fn bpp_vqueue_destroy(q: *mut bpp_vqueue);
    pub fn bpp_vqueue_submit(q: *mut bpp_vqueue, args: *const bpp_verify_args, chunk_status: *mut i32, masks32: *mut u8, mask_present: *mut u8, ticket: *mut u64) -> i32;
    pub fn bpp_vqueue_wait(q: *mut bpp_vqueue, ticket: u64) -> i32;
    pub fn bpp_vqueue_verify(q: *mut bpp_vqueue, args: *const bpp_verify_args, chunk_status: *mut i32, masks32: *mut u8, mask_present: *mut u8) -> i32;
    pub fn bpp_vqueue_stats(q: *mut bpp_vqueue, out5: *mut u64) -> i32;
    pub fn bpp_vqueue_lanes(q: *const bpp_vqueue) -> i32;
    pub fn bpp_vqueue_lane_ms(q: *mut bpp_vqueue, out4: *mut f64) -> i32;
    pub fn bpp_vqueue_set_device_weights(q: *mut bpp_vqueue, enable: i32) -> i32;
    /// one multiscalar check per device pass instead of one per reference call; per-call statuses are unchanged (bpp_b200.h)
    pub fn bpp_vqueue_set_merged_check(q: *mut bpp_vqueue, enable: i32) -> i32;
    pub fn bpp_ctx_set_merged_check(ctx: *mut bpp_ctx, enable: i32) -> i32;
    pub fn bpp_ctx_merged_fallbacks(ctx: *const bpp_ctx) -> u64;
    pub fn bpp_msm_window_bits(n_entries: usize, n_seg: usize) -> i32;
    /// page-locked host memory: proof bytes placed here are uploaded without a staging copy (keep them valid until the call's results are back)
    pub fn bpp_host_alloc(bytes: usize, out: *mut *mut core::ffi::c_void) -> i32;
    pub fn bpp_host_free(p: *mut core::ffi::c_void);

    // batched proving
    pub fn bpp_proof_size(extension_degree: i32, rounds: i32) -> usize;
    pub fn bpp_prove_batch(g: *mut bpp_gens, args: *const bpp_prove_args, proofs_out: *mut u8, proof_stride: usize, status: *mut i32) -> i32;

    // proof bytes, Merlin transcripts on the 203-byte wire form (host)
    pub fn bpp_proof_check_bytes(bytes: *const u8, len: usize, extension_degree: *mut i32, rounds: *mut i32) -> i32;
    pub fn bpp_transcript_new(label: *const u8, len: usize, out: *mut u8);
    pub fn bpp_transcript_append_message(t: *mut u8, label: *const u8, label_len: usize, msg: *const u8, len: usize);
    pub fn bpp_transcript_challenge_bytes(t: *mut u8, label: *const u8, label_len: usize, out: *mut u8, len: usize);
}
