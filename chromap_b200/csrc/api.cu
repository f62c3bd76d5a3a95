// chromap_b200 — C-ABI implementation (include/chromap_b200.h): context, device index / reference,
// batch pipeline driver (tiers, streams, events), stage entry points.  sm_100a.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "../../include/chromap_b200.h"
#include "index_build.cuh"
#include "pipeline_kernels.cuh"
#include "seed_front.cuh"
#include "cta_pair_candidates.cuh"
#include "cta_verify_pairing.cuh"
#include "postprocess.cuh"
#include "exchange.cuh"
#include "sam_kernels.cuh"
#include "ingest.cuh"

static_assert(sizeof(OutRecord) == sizeof(cmx_pe_record), "record layout");
static_assert(sizeof(cmx_pe_record) == 24, "record size");
static_assert(sizeof(OutPairs) == 24 && sizeof(cmx_pairs_record) == 24, "pairs record size");
static_assert(sizeof(OutSam) == sizeof(cmx_sam_record) && sizeof(OutSam) % 4 == 0 && SAM_MAX_CIGAR == CMX_SAM_MAX_CIGAR, "SAM record layout");

#define N_TIERS 3

struct DevBuf {  // grow-only device buffer
  void *p = nullptr;
  size_t cap = 0;
};

struct Tier {
  Caps caps;
  int slots_cap = 0;
  DevBuf mem;
  DevBuf ovf_list;   // pairs that overflowed THIS tier
  int n_slots = 0;   // used in the current batch
  const int *pair_list = nullptr;
  Scratch view;
};

// Per-lane state of the batch pipeline.  A call that carries several whole reference batches is cut into up to
// CMX_MAX_LANES contiguous groups of batches; every group runs the full pipeline on its own stream (one host
// thread each), so the latency-bound kernels of one lane (overflow tiers, candidate pairing) share the SMs with
// the issue-bound kernels of another (minimizers, verification).
#define CMX_MAX_LANES 4
struct Lane {
  DevBuf rescue_list, verify_list, emit_list, nbest, sel, out_rec, out_n, offs, chunk_start, cub_tmp, bc_key, bc_ok, out_compact, bc_out;
  Counters *ctr = nullptr;
  int *d_count = nullptr;
  Tier tiers[N_TIERS];
  cudaStream_t stream = nullptr, aux[N_TIERS - 1] = {nullptr, nullptr};  // aux: emit of the overflow tiers, beside tier 0's
  cudaEvent_t ev_fork = nullptr, ev_join[N_TIERS - 1] = {nullptr, nullptr};
  cudaEvent_t ev[10] = {};
  cudaEvent_t ev_sub[2] = {};
  cudaEvent_t ev_done = nullptr;
  u32 p0 = 0, n = 0;  // pair range of the last call
  int tiers_used = 0;
  std::string err;
};

#define CMX_INGEST_SLOTS 6
struct IngestSlot {  // buffers of one cmx_ingest_fastq stream (text in, packed reads out)
  DevBuf text, nl, seq_start, qual_start, len, off, seq, qual, spans, tmp, stats, count;
  cudaStream_t stream = nullptr;
};

struct cmx_ctx {
  int device = 0;
  cmx_params params;
  DevParams dp;
  std::string err;
  // reference
  u8 *ref_seq = nullptr;
  u64 *ref_off = nullptr;
  u32 *ref_len = nullptr;
  u32 n_seq = 0;
  u64 ref_bytes = 0;
  std::vector<u64> h_ref_off;
  std::vector<u32> h_ref_len;
  // index
  ulonglong2 *slots = nullptr;
  u64 n_slots = 0;
  u64 *occ = nullptr;
  u32 n_occ = 0;
  u64 n_keys = 0;
  int k = 0, w = 0;
  // scATAC barcode whitelist
  ulonglong2 *wl_slots = nullptr;
  u64 wl_n_slots = 0, wl_num_sample = 0;
  double *wl_pow = nullptr;
  u32 wl_bc_len = 0;
  int wl_err = 1, wl_output_nw = 0, wl_active = 0;
  double wl_prob = 0.9;
  DevBuf bc_seq, bc_qual;
  // mapq tables
  double *inv_log = nullptr;
  int *pen_thr = nullptr;
  u32 *mt_init = nullptr;  // std::mt19937(11) right after seeding
  // per-batch buffers
  DevBuf seq1, off1, seq2, off2, trace;
  Lane lanes[CMX_MAX_LANES];
  IngestSlot ingest[CMX_INGEST_SLOTS];
  int sf_grid = 148;            // persistent grid of the front-end kernel: SMs x resident CTAs
  int pcw_grid = 148 * 16;      // persistent grid of the warp-per-pair rescue pass of tier 0
  int n_lanes = CMX_MAX_LANES;  // lanes a multi-batch call is cut into (cmx_set_lanes; CMX_LANES overrides the default)
  int last_lanes_used = 0;
  cudaStream_t stream = nullptr, up_stream = nullptr, down_stream = nullptr;
  // page-locked staging of h2d_big (allocated at the first big upload, kept: page-locking costs more than the copy)
  char *stage[2] = {nullptr, nullptr};
  cudaEvent_t stage_ev[2] = {nullptr, nullptr};
  cudaStream_t stage_stream = nullptr;
  std::vector<cudaEvent_t> ev_up;
  cudaEvent_t ev_bc = nullptr;
  cudaEvent_t ev[4] = {};
  cmx_timing timing;
  u32 last_n_pairs = 0;
  // multi-GPU exchange (cmx_comm_init / cmx_dedup_exchange): an NCCL communicator of this context's own
  void *nccl_comm = nullptr;
  int comm_rank = 0, comm_size = 1;
};

static int fail(cmx_ctx *c, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  return code;
}
#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) return fail(ctx, CMX_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)

static cudaError_t ensure(DevBuf &b, size_t bytes) {
  if (bytes <= b.cap) return cudaSuccess;
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
  const size_t want = bytes + bytes / 8 + 256;
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e == cudaSuccess) b.cap = want;
  return e;
}
static void release(DevBuf &b) { if (b.p) cudaFree(b.p); b.p = nullptr; b.cap = 0; }

extern "C" {

void cmx_default_params(cmx_params *p) {
  p->error_threshold = 8; p->min_num_seeds = 2; p->max_seed_freq0 = 500; p->max_seed_freq1 = 1000;
  p->max_num_best_mappings = 1; p->max_insert_size = 1000; p->mapq_threshold = 30; p->min_read_length = 30;
  p->drop_repetitive_reads = 500000; p->trim_adapters = 0; p->remove_pcr_duplicates = 0; p->tn5_shift = 0;
  p->split_alignment = 0; p->low_memory_mode = 0; p->output_format = 1; p->batch_size = 500000; p->max_read_length = 160;
  p->single_end = 0;
}

int cmx_apply_preset(cmx_params *p, const char *preset) {  // chromap_driver.cc:247-275
  const std::string s = preset ? preset : "";
  if (s.empty()) return CMX_OK;
  if (s == "atac") { p->max_insert_size = 2000; p->trim_adapters = 1; p->remove_pcr_duplicates = 1; p->tn5_shift = 1; p->low_memory_mode = 1; p->output_format = 1; return CMX_OK; }
  if (s == "chip") { p->max_insert_size = 2000; p->remove_pcr_duplicates = 1; p->low_memory_mode = 1; p->output_format = 1; return CMX_OK; }
  if (s == "hic") { p->error_threshold = 4; p->mapq_threshold = 1; p->split_alignment = 1; p->low_memory_mode = 1; p->output_format = 5; return CMX_OK; }
  return CMX_ERR_INVALID;
}

const char *cmx_last_error(const cmx_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int cmx_create(cmx_ctx **out, int device, const cmx_params *params) {
  if (!out || !params) return CMX_ERR_INVALID;
  *out = nullptr;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0 || device >= n_dev) return CMX_ERR_NO_DEVICE;
  if (!(((params->output_format == 1 || params->output_format == 2 || params->output_format == 4) && !params->split_alignment) ||
        (params->output_format == 5 && params->split_alignment)))
    return CMX_ERR_INVALID;  // BED / TagAlign (same records), SAM cores, or Hi-C pairs with split alignment
  if (params->output_format == 4 && (params->max_read_length > SAM_MAX_L || params->error_threshold > SAM_MAX_E)) return CMX_ERR_INVALID;
  if (params->error_threshold < 1 || params->error_threshold >= 16) return CMX_ERR_INVALID;  // mapping_parameters.h:80-88
  if (params->max_num_best_mappings < 1 || params->max_num_best_mappings > CMX_MAX_BEST) return CMX_ERR_INVALID;
  if (params->batch_size < 1 || params->max_read_length < params->min_read_length) return CMX_ERR_INVALID;
  if (params->single_end && (params->split_alignment || params->output_format == 5)) return CMX_ERR_INVALID;  // single-end: BED / TagAlign only
  if (params->output_format == 5 && params->remove_pcr_duplicates && !params->low_memory_mode) return CMX_ERR_INVALID;  // pairs dedup: low-memory rule only
  cmx_ctx *ctx = new cmx_ctx;
  // a failing CUDA call releases what has been built so far (streams, events, device buffers)
#define CUC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cmx_destroy(ctx); return CMX_ERR_CUDA; } } while (0)
  ctx->device = device;
  ctx->params = *params;
  CUC(cudaSetDevice(device));
  // The path's HBM traffic is random 16-byte table slots and short occurrence runs; CMX_L2_FETCH = 32 | 64 | 128 sets the L2
  // fetch-granularity hint for an A/B run (unset: the driver's default — the setting every committed number was measured with).
  if (const char *ev = getenv("CMX_L2_FETCH")) {
    const int gran = atoi(ev);
    if (gran == 32 || gran == 64 || gran == 128) { if (cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)gran) != cudaSuccess) cudaGetLastError(); }
  }
  CUC(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  CUC(cudaStreamCreateWithFlags(&ctx->up_stream, cudaStreamNonBlocking));
  CUC(cudaStreamCreateWithFlags(&ctx->down_stream, cudaStreamNonBlocking));
  for (auto &e : ctx->ev) CUC(cudaEventCreate(&e));
  if (const char *ev = getenv("CMX_LANES")) ctx->n_lanes = std::max(1, std::min(CMX_MAX_LANES, atoi(ev)));
  for (Lane &L : ctx->lanes) {
    CUC(cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking));
    for (auto &a : L.aux) CUC(cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking));
    CUC(cudaEventCreateWithFlags(&L.ev_fork, cudaEventDisableTiming));
    for (auto &e : L.ev_join) CUC(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto &e : L.ev) CUC(cudaEventCreate(&e));
    for (auto &e : L.ev_sub) CUC(cudaEventCreate(&e));
    CUC(cudaEventCreateWithFlags(&L.ev_done, cudaEventDisableTiming));
    CUC(cudaMalloc(&L.ctr, sizeof(Counters)));
    CUC(cudaMalloc(&L.d_count, sizeof(int) * 4));
  }
  // MAPQ tables from the host libm, so truncations match the reference bit for bit (mapping_generator.h:920-1022)
  {
    std::vector<double> il(65536, 0.0);
    const int coef_frac = log(50);
    for (int a = 2; a < 65536; ++a) il[a] = coef_frac / log((unsigned short)a);
    std::vector<int> thr(96, 0x7fffffff);
    for (int v = 0; v < 96; ++v) {  // smallest n >= 0 with (int)(4.343*log(n+1)+0.499) >= v (monotone in n)
      long long lo = 0, hi = 0x7ffffffeLL;
      auto pen = [](long long n) { return (int)(4.343 * log((double)(n + 1)) + 0.499); };
      if (pen(hi) < v) { thr[v] = 0x7fffffff; continue; }
      while (lo < hi) { const long long mid = (lo + hi) / 2; if (pen(mid) >= v) hi = mid; else lo = mid + 1; }
      thr[v] = (int)lo;
    }
    CUC(cudaMalloc(&ctx->inv_log, 65536 * sizeof(double)));
    CUC(cudaMalloc(&ctx->pen_thr, 96 * sizeof(int)));
    CUC(cudaMemcpy(ctx->inv_log, il.data(), 65536 * sizeof(double), cudaMemcpyHostToDevice));
    CUC(cudaMemcpy(ctx->pen_thr, thr.data(), 96 * sizeof(int), cudaMemcpyHostToDevice));
  }
  {
    std::vector<u32> mt(624);
    mt[0] = 11u;
    for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (u32)i;
    CUC(cudaMalloc(&ctx->mt_init, 624 * sizeof(u32)));
    CUC(cudaMemcpy(ctx->mt_init, mt.data(), 624 * sizeof(u32), cudaMemcpyHostToDevice));
  }
  // the overflow-tier kernels may use more than the default 48 KB of (static + dynamic) shared memory
  CUC(cudaFuncSetAttribute(cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * CLUSTER_NT * 8));  // tier-0 hc = 64
  CUC(cudaFuncSetAttribute(seed_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  CUC(cudaFuncSetAttribute(pair_candidates_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  CUC(cudaFuncSetAttribute(verify_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  CUC(cudaFuncSetAttribute(pairing_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  CUC(cudaFuncSetAttribute(verify_split_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  const int mrl = params->max_read_length;
  if ((size_t)2 * mrl * 64 > 48 * 1024) {  // per-thread read-code columns of the verification kernels (long reads)
    if ((size_t)2 * mrl * 64 > 200 * 1024) { cmx_destroy(ctx); return CMX_ERR_INVALID; }
    CUC(cudaFuncSetAttribute(verify_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * mrl * 64));
  }
  {
    const size_t sf_smem = seed_front_smem_bytes(mrl);
    CUC(cudaFuncSetAttribute(seed_front_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sf_smem));
    CUC(cudaFuncSetAttribute(seed_front_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sf_smem));
    int per_sm = 0, n_sm = 0;
    CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, seed_front_kernel<true>, SF_NT, sf_smem));
    CUC(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
    ctx->sf_grid = std::max(1, per_sm) * std::max(1, n_sm);
    int per_sm_w = 0;
    CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_w, pair_candidates_cta_kernel, 32, pair_candidates_cta_smem(64, 32, mrl, 64)));
    ctx->pcw_grid = std::max(1, per_sm_w) * std::max(1, n_sm);
  }
  for (Lane &L : ctx->lanes) {
    L.tiers[0].caps = {mrl, 64, 32, 32};
    L.tiers[1].caps = {mrl * 2, 1024, 256, 256};
    L.tiers[2].caps = {mrl * 4, 65536, 8192, 8192};
  }
  memset(&ctx->timing, 0, sizeof(ctx->timing));
#undef CUC
  *out = ctx;
  return CMX_OK;
}

void cmx_destroy(cmx_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cmx_comm_destroy(ctx);
  cudaFree(ctx->ref_seq); cudaFree(ctx->ref_off); cudaFree(ctx->ref_len);
  cudaFree(ctx->slots); cudaFree(ctx->occ); cudaFree(ctx->inv_log); cudaFree(ctx->pen_thr); cudaFree(ctx->mt_init);
  cudaFree(ctx->wl_slots); cudaFree(ctx->wl_pow);
  for (DevBuf *b : {&ctx->bc_seq, &ctx->bc_qual, &ctx->seq1, &ctx->off1, &ctx->seq2, &ctx->off2, &ctx->trace}) release(*b);
  for (Lane &L : ctx->lanes) {
    cudaFree(L.ctr); cudaFree(L.d_count);
    for (DevBuf *b : {&L.rescue_list, &L.verify_list, &L.emit_list, &L.nbest, &L.sel, &L.out_rec, &L.out_n, &L.offs, &L.chunk_start, &L.cub_tmp, &L.bc_key, &L.bc_ok, &L.out_compact, &L.bc_out})
      release(*b);
    for (auto &t : L.tiers) { release(t.mem); release(t.ovf_list); }
    for (auto &e : L.ev) if (e) cudaEventDestroy(e);
    for (auto &e : L.ev_sub) if (e) cudaEventDestroy(e);
    if (L.ev_done) cudaEventDestroy(L.ev_done);
    if (L.stream) cudaStreamDestroy(L.stream);
    for (auto &a : L.aux) if (a) cudaStreamDestroy(a);
    if (L.ev_fork) cudaEventDestroy(L.ev_fork);
    for (auto &e : L.ev_join) if (e) cudaEventDestroy(e);
  }
  for (IngestSlot &g : ctx->ingest) {
    for (DevBuf *b : {&g.text, &g.nl, &g.seq_start, &g.qual_start, &g.len, &g.off, &g.seq, &g.qual, &g.spans, &g.tmp, &g.stats, &g.count}) release(*b);
    if (g.stream) cudaStreamDestroy(g.stream);
  }
  for (auto &e : ctx->ev) if (e) cudaEventDestroy(e);
  for (auto &e : ctx->ev_up) cudaEventDestroy(e);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->up_stream) cudaStreamDestroy(ctx->up_stream);
  if (ctx->stage[0]) cudaFreeHost(ctx->stage[0]);
  if (ctx->stage[1]) cudaFreeHost(ctx->stage[1]);
  for (auto &e : ctx->stage_ev) if (e) cudaEventDestroy(e);
  if (ctx->stage_stream) cudaStreamDestroy(ctx->stage_stream);
  if (ctx->down_stream) cudaStreamDestroy(ctx->down_stream);
  delete ctx;
}

// Pageable or file-mapped host memory -> device faster than one thread can copy it into the driver's staging area: eight
// threads fill one of two page-locked 32 MB buffers while cudaMemcpyAsync drains the other (a 3 Gbp index is 18 GB of
// arrays that the CLI hands over where they lie in the page cache).  Pinned or device sources go straight through.
static const size_t STAGE_BYTES = 32u << 20;
static cudaError_t h2d_big(cmx_ctx *ctx, void *dst, const void *src, size_t bytes) {
  if (bytes == 0) return cudaSuccess;
  cudaPointerAttributes at;
  const bool plain = cudaPointerGetAttributes(&at, src) == cudaSuccess && at.type == cudaMemoryTypeUnregistered;
  cudaGetLastError();
  if (!plain || bytes < STAGE_BYTES / 2) return cudaMemcpy(dst, src, bytes, cudaMemcpyDefault);
  if (!ctx->stage_stream) {
    if (cudaMallocHost(&ctx->stage[0], STAGE_BYTES) != cudaSuccess || cudaMallocHost(&ctx->stage[1], STAGE_BYTES) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->stage_ev[0], cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&ctx->stage_ev[1], cudaEventDisableTiming) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->stage_stream, cudaStreamNonBlocking) != cudaSuccess) {
      cudaGetLastError();
      ctx->stage_stream = nullptr;
      return cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
    }
  }
  cudaStream_t st = ctx->stage_stream;
  cudaError_t rc = cudaSuccess;
  size_t i = 0;
  for (size_t o = 0; o < bytes && rc == cudaSuccess; o += STAGE_BYTES, ++i) {
    const int b = (int)(i & 1);
    const size_t n = std::min(STAGE_BYTES, bytes - o);
    if (i >= 2) rc = cudaEventSynchronize(ctx->stage_ev[b]);
    if (rc != cudaSuccess) break;
    const int nt = 8;
    const size_t slice = ((n + nt - 1) / nt + 4095) & ~(size_t)4095;
    char *stage = ctx->stage[b];
    std::thread th[8];
    for (int t = 0; t < nt; ++t)
      th[t] = std::thread([=]() {
        const size_t a = std::min(n, (size_t)t * slice), z = std::min(n, a + slice);
        if (z > a) memcpy(stage + a, (const char *)src + o + a, z - a);
      });
    for (int t = 0; t < nt; ++t) th[t].join();
    rc = cudaMemcpyAsync((char *)dst + o, stage, n, cudaMemcpyHostToDevice, st);
    if (rc == cudaSuccess) rc = cudaEventRecord(ctx->stage_ev[b], st);
  }
  const cudaError_t rs = cudaStreamSynchronize(st);
  return rc == cudaSuccess ? rs : rc;
}

int cmx_upload_reference(cmx_ctx *ctx, uint32_t n_seq, const uint64_t *offsets, const char *concat) {
  if (!ctx || !offsets || !concat || n_seq == 0) return CMX_ERR_INVALID;
  CU(cudaSetDevice(ctx->device));
  cudaFree(ctx->ref_seq); cudaFree(ctx->ref_off); cudaFree(ctx->ref_len);
  ctx->ref_seq = nullptr; ctx->ref_off = nullptr; ctx->ref_len = nullptr;
  // device layout: [64 NUL][seq0][64.. NUL][seq1]...  every sequence start 64-byte aligned, >= 64 NULs after each
  const u64 PAD = 64;
  std::vector<u64> doff(n_seq);
  std::vector<u32> dlen(n_seq);
  u64 cur = PAD;
  for (u32 i = 0; i < n_seq; ++i) {
    const u64 len = offsets[i + 1] - offsets[i];
    if (len >= 0xFFFFFFFFull) return fail(ctx, CMX_ERR_INVALID, "reference sequence %u too long for 32-bit positions", i);
    doff[i] = cur; dlen[i] = (u32)len;
    cur = (cur + len + PAD + 63) / 64 * 64;
  }
  ctx->ref_bytes = cur + PAD;
  CU(cudaMalloc(&ctx->ref_seq, ctx->ref_bytes));
  CU(cudaMemset(ctx->ref_seq, 0, ctx->ref_bytes));
  CU(cudaDeviceSynchronize());  // (the staged copies below run on their own stream)
  for (u32 i = 0; i < n_seq; ++i)
    CU(h2d_big(ctx, ctx->ref_seq + doff[i], concat + offsets[i], dlen[i]));  // host or device source
  CU(cudaMalloc(&ctx->ref_off, n_seq * sizeof(u64)));
  CU(cudaMalloc(&ctx->ref_len, n_seq * sizeof(u32)));
  CU(cudaMemcpy(ctx->ref_off, doff.data(), n_seq * sizeof(u64), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(ctx->ref_len, dlen.data(), n_seq * sizeof(u32), cudaMemcpyHostToDevice));
  ctx->n_seq = n_seq; ctx->h_ref_off = doff; ctx->h_ref_len = dlen;
  return CMX_OK;
}

// Allocate the device table for n_keys keys (load <= 0.5) and clear it.
static int alloc_table(cmx_ctx *ctx, u64 n_keys) {
  cudaFree(ctx->slots); ctx->slots = nullptr;
  u64 n = 1024;
  while (n < 2 * n_keys) n <<= 1;
  ctx->n_slots = n;
  CU(cudaMalloc(&ctx->slots, n * sizeof(ulonglong2)));
  CU(cudaMemset(ctx->slots, 0xFF, n * sizeof(ulonglong2)));
  return CMX_OK;
}
static int table_shift(u64 n_slots) { int lg = 0; while ((1ull << lg) < n_slots) ++lg; return 64 - lg; }

// The mate-guided lookup (cta_pair_candidates.cuh) relies on every occurrence list holding distinct reference positions
// (true for any index Index::Construct builds: one k-mer per position).  Adjacent equal positions would break it: refuse.
__global__ void occ_check_kernel(const u64 *occ, u32 n, unsigned long long *bad) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i + 1 < n && (occ[i] >> 1) == (occ[i + 1] >> 1)) atomicAdd(bad, 1ull);
}
static int check_occurrences(cmx_ctx *ctx) {
  if (ctx->n_occ < 2) return CMX_OK;
  unsigned long long *d_bad = nullptr, bad = 0;
  CU(cudaMalloc(&d_bad, 8));
  CU(cudaMemset(d_bad, 0, 8));
  occ_check_kernel<<<(ctx->n_occ + 255) / 256, 256>>>(ctx->occ, ctx->n_occ, d_bad);
  CU(cudaMemcpy(&bad, d_bad, 8, cudaMemcpyDeviceToHost));
  cudaFree(d_bad);
  if (bad) return fail(ctx, CMX_ERR_INVALID, "index: %llu adjacent occurrence entries share a reference position (not an index Chromap builds)", bad);
  return CMX_OK;
}

int cmx_upload_index(cmx_ctx *ctx, int k, int w, uint32_t n_buckets, const uint32_t *flags, const uint64_t *keys,
                     const uint64_t *vals, const uint64_t *occ, uint32_t n_occ) {
  if (!ctx || !flags || !keys || !vals || (n_occ && !occ)) return CMX_ERR_INVALID;
  if (k < 2 || k > 28 || w < 1 || w > CMX_W_MAX) return fail(ctx, CMX_ERR_INVALID, "unsupported k=%d w=%d", k, w);
  CU(cudaSetDevice(ctx->device));
  // occupied buckets (khash.h:165: 2 flag bits per bucket: 10 empty, 01 deleted, 00 occupied) are counted and inserted on
  // the device, straight from the reference's arrays: no host-side compaction pass over a billion buckets
  const u64 nbk = n_buckets;
  const size_t nfw = (size_t)((nbk + 15) / 16);
  u32 *d_flags = nullptr;
  unsigned long long *d_cnt = nullptr;
  u64 *d_k = nullptr, *d_v = nullptr;
  CU(cudaMalloc(&d_flags, nfw * 4)); CU(cudaMalloc(&d_cnt, 8));
  CU(h2d_big(ctx, d_flags, flags, nfw * 4));
  CU(cudaMemset(d_cnt, 0, 8));
  khash_count_kernel<<<(unsigned)((nfw + 255) / 256), 256>>>(d_flags, nbk, d_cnt);
  unsigned long long n_keys = 0;
  CU(cudaMemcpy(&n_keys, d_cnt, 8, cudaMemcpyDeviceToHost));
  ctx->n_keys = n_keys;
  int rc = alloc_table(ctx, n_keys);
  if (rc) return rc;
  const size_t CH = 1u << 25;  // buckets per chunk: 2 x 256 MB in flight
  CU(cudaMalloc(&d_k, std::min<size_t>(nbk, CH) * 8 + 16)); CU(cudaMalloc(&d_v, std::min<size_t>(nbk, CH) * 8 + 16));
  for (u64 b0 = 0; b0 < nbk; b0 += CH) {
    const u64 n = std::min<u64>(CH, nbk - b0);
    CU(cudaStreamSynchronize(0));  // the previous chunk's insert kernel is done with d_k / d_v (the staged copies run on their own stream)
    CU(h2d_big(ctx, d_k, keys + b0, n * 8));
    CU(h2d_big(ctx, d_v, vals + b0, n * 8));
    khash_insert_kernel<<<(unsigned)((n + 255) / 256), 256>>>(d_flags, d_k, d_v, b0, n, ctx->slots, ctx->n_slots - 1, table_shift(ctx->n_slots));
    CU(cudaGetLastError());
  }
  CU(cudaDeviceSynchronize());
  cudaFree(d_flags); cudaFree(d_cnt); cudaFree(d_k); cudaFree(d_v);
  cudaFree(ctx->occ); ctx->occ = nullptr;
  CU(cudaMalloc(&ctx->occ, (size_t)std::max<u32>(n_occ, 1) * sizeof(u64)));
  if (n_occ) CU(h2d_big(ctx, ctx->occ, occ, (size_t)n_occ * sizeof(u64)));
  ctx->n_occ = n_occ; ctx->k = k; ctx->w = w;
  return check_occurrences(ctx);
}

int cmx_upload_barcode_whitelist(cmx_ctx *ctx, const uint64_t *keys, const uint32_t *counts, uint64_t n, uint64_t num_sample, uint32_t bc_len,
                                 int err_threshold, double prob_threshold, int output_not_in_whitelist) {
  if (!ctx || (n && (!keys || !counts)) || bc_len == 0 || bc_len > 32) return CMX_ERR_INVALID;
  if (err_threshold < 0 || err_threshold > 1) return fail(ctx, CMX_ERR_INVALID, "--bc-error-threshold %d is not on the GPU path (0 or 1)", err_threshold);
  CU(cudaSetDevice(ctx->device));
  cudaFree(ctx->wl_slots); ctx->wl_slots = nullptr;
  u64 ns = 64;
  while (ns < 2 * n) ns <<= 1;
  CU(cudaMalloc(&ctx->wl_slots, ns * sizeof(ulonglong2)));
  CU(cudaMemset(ctx->wl_slots, 0xFF, ns * sizeof(ulonglong2)));
  if (n) {
    u64 *dk; u32 *dc;
    CU(cudaMalloc(&dk, n * 8)); CU(cudaMalloc(&dc, n * 4));
    CU(cudaMemcpy(dk, keys, n * 8, cudaMemcpyHostToDevice)); CU(cudaMemcpy(dc, counts, n * 4, cudaMemcpyHostToDevice));
    wl_insert_kernel<<<(unsigned)((n + 255) / 256), 256>>>(dk, dc, n, ctx->wl_slots, ns - 1, table_shift(ns));
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
    cudaFree(dk); cudaFree(dc);
  }
  if (!ctx->wl_pow) {
    std::vector<double> pw(41);
    for (int q = 0; q <= 40; ++q) pw[q] = pow(10.0, ((-q) / 10.0));  // host libm, chromap.cc:629-630
    CU(cudaMalloc(&ctx->wl_pow, 41 * sizeof(double)));
    CU(cudaMemcpy(ctx->wl_pow, pw.data(), 41 * sizeof(double), cudaMemcpyHostToDevice));
  }
  ctx->wl_n_slots = ns; ctx->wl_num_sample = num_sample; ctx->wl_bc_len = bc_len; ctx->wl_err = err_threshold; ctx->wl_prob = prob_threshold;
  ctx->wl_output_nw = output_not_in_whitelist; ctx->wl_active = 1;
  return CMX_OK;
}

int cmx_index_info(const cmx_ctx *ctx, int *k, int *w, uint64_t *n_keys, uint64_t *n_occ, uint64_t *table_slots) {
  if (!ctx || !ctx->slots) return CMX_ERR_STATE;
  if (k) *k = ctx->k;
  if (w) *w = ctx->w;
  if (n_keys) *n_keys = ctx->n_keys;
  if (n_occ) *n_occ = ctx->n_occ;
  if (table_slots) *table_slots = ctx->n_slots;
  return CMX_OK;
}

int cmx_build_index(cmx_ctx *ctx, int k, int w) {
  if (!ctx) return CMX_ERR_INVALID;
  if (!ctx->ref_seq) return fail(ctx, CMX_ERR_STATE, "upload the reference first");
  if (k < 2 || k > 28 || w < 1 || w > CMX_W_MAX) return fail(ctx, CMX_ERR_INVALID, "unsupported k=%d w=%d", k, w);
  CU(cudaSetDevice(ctx->device));
  IndexBuildResult r;
  std::string err;
  const int rc = build_index_on_device(ctx->ref_seq, ctx->h_ref_off, ctx->h_ref_len, k, w, &r, &err);
  if (rc != 0) return fail(ctx, rc, "%s", err.c_str());
  cudaFree(ctx->slots); cudaFree(ctx->occ);
  ctx->slots = r.slots; ctx->n_slots = r.n_slots; ctx->occ = r.occ; ctx->n_occ = r.n_occ; ctx->n_keys = r.n_keys;
  ctx->k = k; ctx->w = w;
  return check_occurrences(ctx);
}

// khash geometry + layout on the device: every (key, val) re-inserted with khash's own probe sequence
// (hash = key>>1 truncated to 32 bit, triangular steps; khash.h:232-245) so the reference's kh_get finds it.
__global__ void khash_layout_kernel(const ulonglong2 *slots, u64 n_slots, u64 *keys, u64 *vals, u32 mask) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_slots) return;
  const ulonglong2 kv = slots[i];
  if (kv.x == CMX_EMPTY_KEY) return;
  u32 b = (u32)(kv.x >> 1) & mask, step = 0;
  for (;;) {
    const u64 old = atomicCAS((unsigned long long *)&keys[b], (unsigned long long)CMX_EMPTY_KEY, (unsigned long long)kv.x);
    if (old == CMX_EMPTY_KEY) { vals[b] = kv.y; return; }
    b = (b + (++step)) & mask;
  }
}
__global__ void khash_flags_kernel(u64 *keys, u64 *vals, u32 n_buckets, u32 *flags) {
  const u32 wi = blockIdx.x * blockDim.x + threadIdx.x;  // one flag word = 16 buckets (khash.h:165)
  const u32 nf = n_buckets < 16 ? 1 : n_buckets >> 4;
  if (wi >= nf) return;
  u32 f = 0;
  for (u32 j = 0; j < 16; ++j) {
    const u32 b = wi * 16 + j;
    const bool empty = b >= n_buckets || keys[b] == CMX_EMPTY_KEY;
    if (empty) { f |= 2u << (j << 1); if (b < n_buckets) { keys[b] = 0; vals[b] = 0; } }
  }
  flags[wi] = f;
}

int cmx_download_index(cmx_ctx *ctx, uint32_t *n_buckets, uint32_t *n_keys, uint32_t *flags, uint64_t *keys, uint64_t *vals,
                       uint32_t *n_occ, uint64_t *occ) {
  if (!ctx || !ctx->slots) return CMX_ERR_STATE;
  CU(cudaSetDevice(ctx->device));
  // bucket count the reference's khash has after inserting n_keys keys (khash.h:295-300: it grows when
  // n_occupied >= upper_bound = n_buckets*0.77+0.5)
  u32 nb = 4;
  while ((u32)(nb * 0.77 + 0.5) < ctx->n_keys) nb <<= 1;
  if (n_buckets) *n_buckets = nb;
  if (n_keys) *n_keys = (u32)ctx->n_keys;
  if (n_occ) *n_occ = ctx->n_occ;
  if (occ && ctx->n_occ) CU(cudaMemcpy(occ, ctx->occ, (size_t)ctx->n_occ * sizeof(u64), cudaMemcpyDeviceToHost));
  if (flags && keys && vals) {
    const size_t nf = nb < 16 ? 1 : nb >> 4;
    u64 *dk = nullptr, *dv = nullptr;
    u32 *df = nullptr;
    CU(cudaMalloc(&dk, (size_t)nb * 8)); CU(cudaMalloc(&dv, (size_t)nb * 8)); CU(cudaMalloc(&df, nf * 4));
    CU(cudaMemset(dk, 0xFF, (size_t)nb * 8)); CU(cudaMemset(dv, 0, (size_t)nb * 8));
    khash_layout_kernel<<<(unsigned)((ctx->n_slots + 255) / 256), 256>>>(ctx->slots, ctx->n_slots, dk, dv, nb - 1);
    khash_flags_kernel<<<(unsigned)((nf + 255) / 256), 256>>>(dk, dv, nb, df);
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(keys, dk, (size_t)nb * 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(vals, dv, (size_t)nb * 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(flags, df, nf * 4, cudaMemcpyDeviceToHost));
    cudaFree(dk); cudaFree(dv); cudaFree(df);
  }
  return CMX_OK;
}

// ---------------------------------------------------------------------------------------------------
static size_t tier_bytes(const Caps &c, size_t slots, bool interleaved, size_t *o) {  // o[11]
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t r = off; off = (off + bytes + 255) / 256 * 256; return r; };
  const size_t R = 2 * slots;
  const size_t Rm = interleaved ? 2 * ((slots + 31) / 32 * 32) : R;  // minimizer records: whole groups of 32 pairs
  o[0] = take(R * sizeof(ReadMeta));
  o[1] = take(slots * sizeof(PairMeta));
  o[2] = take(interleaved ? 0 : R * c.maxmm * 8);  // hashes: tier 0 never stores them
  o[3] = take(Rm * c.maxmm * 8);
  o[4] = take(Rm * c.maxmm * 4);
  o[5] = take(R * 2 * (size_t)c.hc * 8);
  o[6] = take(R * 6 * (size_t)c.cc * 8);
  o[7] = take(R * 6 * (size_t)c.cc);
  o[8] = take(R * 2 * (size_t)c.mc * 8);
  o[9] = take(R * 2 * (size_t)c.mc * 2);
  o[10] = take(R * 2 * (size_t)c.mc * 4);
  return off;
}
static cudaError_t tier_prepare(Tier &t, int n_slots, const int *pair_list, bool interleaved) {
  if (n_slots > t.slots_cap) {
    size_t o[11];
    const size_t want_slots = (size_t)n_slots + n_slots / 8 + 16;
    const size_t bytes = tier_bytes(t.caps, want_slots, interleaved, o);
    release(t.mem);
    cudaError_t e = ensure(t.mem, bytes);
    if (e != cudaSuccess) return e;
    t.slots_cap = (int)want_slots;
  }
  size_t o[11];
  tier_bytes(t.caps, t.slots_cap, interleaved, o);
  char *b = (char *)t.mem.p;
  Scratch &S = t.view;
  S.caps = t.caps; S.n_slots = n_slots; S.pair_list = pair_list; S.mm_il = interleaved ? 1 : 0;
  S.rmeta = (ReadMeta *)(b + o[0]); S.pmeta = (PairMeta *)(b + o[1]);
  S.mm_hash = (u64 *)(b + o[2]); S.mm_val = (u64 *)(b + o[3]); S.mm_pos = (u32 *)(b + o[4]);
  S.hits = (u64 *)(b + o[5]); S.cand_pos = (u64 *)(b + o[6]); S.cand_cnt = (u8 *)(b + o[7]);
  S.map_pos = (u64 *)(b + o[8]); S.map_err = (short *)(b + o[9]); S.map_split = (int *)(b + o[10]);
  t.n_slots = n_slots; t.pair_list = pair_list;
  return cudaSuccess;
}

// Task chunks of `#pragma omp taskloop grainsize(5000)` (chromap.h:892) over one reference batch of n pairs
// as cut by libgomp: num_tasks = n/5000 (min 1), chunk = n/num_tasks, first n%num_tasks chunks one longer.
static void taskloop_chunks(u32 base, u32 n, std::vector<int> &starts) {
  u32 nt = n / 5000;
  if (nt < 1) nt = 1;
  const u32 chunk = n / nt, rem = n % nt;
  u32 s = base;
  for (u32 t = 0; t < nt; ++t) { starts.push_back((int)s); s += chunk + (t < rem ? 1 : 0); }
}

static DevParams make_dev_params(const cmx_ctx *ctx) {
  const cmx_params &p = ctx->params;
  DevParams d;
  d.e = p.error_threshold; d.min_seeds = p.min_num_seeds; d.f0 = p.max_seed_freq0; d.f1 = p.max_seed_freq1;
  d.max_best = p.max_num_best_mappings; d.max_insert = p.max_insert_size; d.min_read_len = p.min_read_length;
  d.drop_rep = p.drop_repetitive_reads; d.trim = p.trim_adapters; d.k = ctx->k; d.w = ctx->w;
  d.lanes = p.error_threshold < 8 ? 8 : 4;  // mapping_parameters.h:80-88
  d.split = p.split_alignment;
  d.se = p.single_end;
  return d;
}

static int upload_batch(cmx_ctx *ctx, const cmx_batch *in, DevBatch *B) {
  const u32 n = in->n_pairs;
  if (in->on_device) {
    B->seq1 = (const u8 *)in->seq1; B->off1 = in->off1; B->seq2 = (const u8 *)in->seq2; B->off2 = in->off2;
  } else {
    const size_t b1 = in->off1[n], b2 = in->off2[n];
    CU(ensure(ctx->seq1, b1 + 64)); CU(ensure(ctx->seq2, b2 + 64));
    CU(ensure(ctx->off1, (n + 1) * 4)); CU(ensure(ctx->off2, (n + 1) * 4));
    CU(cudaMemcpyAsync(ctx->seq1.p, in->seq1, b1, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->seq2.p, in->seq2, b2, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->off1.p, in->off1, (n + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->off2.p, in->off2, (n + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    B->seq1 = (const u8 *)ctx->seq1.p; B->off1 = (const u32 *)ctx->off1.p;
    B->seq2 = (const u8 *)ctx->seq2.p; B->off2 = (const u32 *)ctx->off2.p;
  }
  B->n_pairs = n; B->first_read_id = in->first_read_id; B->bc_ok = nullptr;
  return CMX_OK;
}

struct BatchAcc {  // per-call accumulators over sub-batches
  float ms_seed = 0, ms_pc = 0, ms_ver = 0, ms_pair = 0, ms_minimizer = 0, ms_probe = 0, ms_cluster = 0, ms_select = 0, ms_emit = 0;
  u64 launches = 0, n_overflow = 0;
  Counters c;
  u64 tier_pairs[N_TIERS] = {0, 0, 0};
  int tiers_used = 0;
  BatchAcc() { memset(&c, 0, sizeof(c)); }
};

#define CUL(call)                                                                                  \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      char b_[512];                                                                                \
      snprintf(b_, sizeof(b_), "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      L.err = b_;                                                                                  \
      return CMX_ERR_CUDA;                                                                         \
    }                                                                                              \
  } while (0)

struct LaneJob {  // what one lane maps in one call: pairs [p0, p0 + n) of the call's batch
  u32 p0 = 0, n = 0;
  u32 piece = 0, piece0 = 0;                       // reads arrive in pieces of `piece` pairs; this lane's first piece index
  const std::vector<cudaEvent_t> *piece_ready = nullptr;
  const u8 *bc_seq = nullptr, *bc_qual = nullptr;  // device, already offset to p0
  u32 bc_len = 0;
  cudaEvent_t bc_ready = nullptr;
  bool want_bc = false;
  OutRecord *dst = nullptr;                        // device: compacted records of this lane (nullptr = lane buffer)
  u64 total = 0;
  BatchAcc acc;
  int rc = CMX_OK;
};

// The whole device pipeline over pairs already resident in HBM (or landing piece by piece); records compacted in
// read order into the lane's buffer.  Synchronous on the lane's stream; safe to run one lane per host thread.
static int run_lane(cmx_ctx *ctx, Lane &L, const DevBatch &Bfull, LaneJob &J) {
  CUL(cudaSetDevice(ctx->device));
  const u32 n = J.n;
  BatchAcc &acc = J.acc;
  DevBatch B = Bfull;  // view of this lane's pairs: pair index i of the lane = pair p0 + i of the call
  B.off1 += J.p0; B.off2 += J.p0; B.n_pairs = n; B.first_read_id = Bfull.first_read_id + J.p0; B.bc_ok = nullptr;
  L.p0 = J.p0; L.n = n;
  const int mb = ctx->params.max_num_best_mappings;
  cudaStream_t st = L.stream;
  const DevParams P = make_dev_params(ctx);
  DevIndex ix;
  ix.slots = ctx->slots; ix.n_slots_mask = ctx->n_slots - 1; ix.shift = table_shift(ctx->n_slots); ix.occ = ctx->occ; ix.n_occ = ctx->n_occ; ix.k = ctx->k; ix.w = ctx->w;
  DevRef R;
  R.seq = ctx->ref_seq; R.off = ctx->ref_off; R.len = ctx->ref_len; R.n_seq = ctx->n_seq;
  MapqTables T;
  T.inv_log = ctx->inv_log; T.pen_thr = ctx->pen_thr;
  CUL(ensure(L.nbest, (size_t)n * 4)); CUL(ensure(L.sel, (size_t)n * mb * 4));
  const bool sam = ctx->params.output_format == 4;
  const size_t rec_bytes = sam ? sizeof(OutSam) : sizeof(OutRecord);
  CUL(ensure(L.out_rec, (size_t)n * mb * rec_bytes)); CUL(ensure(L.out_n, (size_t)(n + 1) * 4));
  CUL(ensure(L.offs, (size_t)(n + 1) * 8));
  OutRecord *dst = J.dst;
  if (!dst) { CUL(ensure(L.out_compact, (size_t)n * mb * rec_bytes)); dst = (OutRecord *)L.out_compact.p; }
  J.dst = dst;
  u64 *bc_dst = nullptr;
  if (J.want_bc && J.bc_seq) { CUL(ensure(L.bc_out, (size_t)n * mb * 8)); bc_dst = (u64 *)L.bc_out.p; }
  CUL(cudaMemsetAsync(L.nbest.p, 0, (size_t)n * 4, st));
  CUL(cudaMemsetAsync(L.out_n.p, 0, (size_t)(n + 1) * 4, st));
  CUL(cudaMemsetAsync(L.ctr, 0, sizeof(Counters), st));
  if (J.bc_seq) {  // the barcode gate, when barcodes came with the batch
    DevWhitelist W;
    W.slots = ctx->wl_slots; W.mask = ctx->wl_n_slots ? ctx->wl_n_slots - 1 : 0; W.shift = ctx->wl_n_slots ? table_shift(ctx->wl_n_slots) : 0;
    W.num_sample = (double)ctx->wl_num_sample; W.pow_tab = ctx->wl_pow; W.err_threshold = ctx->wl_err; W.prob_threshold = ctx->wl_prob;
    W.output_not_in_whitelist = ctx->wl_output_nw; W.active = ctx->wl_active;
    CUL(ensure(L.bc_key, (size_t)n * 8)); CUL(ensure(L.bc_ok, (size_t)n));
    if (J.bc_ready) CUL(cudaStreamWaitEvent(st, J.bc_ready, 0));
    barcode_kernel<<<(n + 127) / 128, 128, 0, st>>>(W, J.bc_seq, J.bc_qual, (int)J.bc_len, (int)n, (u64 *)L.bc_key.p, (u8 *)L.bc_ok.p, L.ctr);
    acc.launches += 1;
    B.bc_ok = (const u8 *)L.bc_ok.p;
  }
  int n_slots = (int)n;
  const int *pair_list = nullptr;
  int tiers_used = 0;
  u64 n_overflow_final = 0;
  const int TB = 128;
  for (int t = 0; t < N_TIERS && n_slots > 0; ++t) {
    Tier &tier = L.tiers[t];
    CUL(tier_prepare(tier, n_slots, pair_list, t == 0));
    const Scratch S = tier.view;
    cudaEvent_t e0 = L.ev[5], e1 = L.ev[6], e2 = L.ev[7], e3 = L.ev[8], e4 = L.ev[9];
    int cluster_passes = 1;
    CUL(cudaEventRecord(e0, st));
    if (t == 0) {
      // front end: [adapter trimming] + length filter + minimizers + index probe in one kernel over staged read tiles
      // (seed_front.cuh).  When the reads arrive in pieces on the upload stream, one grid per piece starts as soon as
      // the piece has landed; the rest of the upload hides behind it.
      const size_t sf_smem = seed_front_smem_bytes(S.caps.maxmm);
      const u32 piece = J.piece_ready ? J.piece : n;
      for (u32 q = 0, p0 = 0; p0 < n; ++q, p0 += piece) {
        const u32 np = std::min(piece, n - p0);
        if (J.piece_ready) CUL(cudaStreamWaitEvent(st, (*J.piece_ready)[J.piece0 + q], 0));
        if (P.trim) {
          Scratch V = S;
          V.n_slots = (int)np; V.rmeta += 2 * (size_t)p0; V.pmeta += p0;
          DevBatch Bq = B;
          if (Bq.bc_ok) Bq.bc_ok += p0;
          Bq.off1 += p0; Bq.off2 += p0; Bq.n_pairs = np;
          prep_kernel<<<(np + TB - 1) / TB, TB, 0, st>>>(P, Bq, V);
          acc.launches += 1;
        }
        const int tiles = (int)((np + SF_TILE - 1) / SF_TILE);
        // packed-key scan for k = 17, w = 7 and reads the key layout can address (minimizers.cuh); the run-time scan otherwise
        if (P.k == 17 && P.w == 7 && S.caps.maxmm < (1 << 18))
          seed_front_kernel<true><<<std::min(tiles, ctx->sf_grid), SF_NT, sf_smem, st>>>(P, ix, B, S, L.ctr, P.trim ? 1 : 0, (int)p0, (int)(p0 + np));
        else
          seed_front_kernel<false><<<std::min(tiles, ctx->sf_grid), SF_NT, sf_smem, st>>>(P, ix, B, S, L.ctr, P.trim ? 1 : 0, (int)p0, (int)(p0 + np));
        acc.launches += 1;
      }
    } else {
      prep_kernel<<<(n_slots + TB - 1) / TB, TB, 0, st>>>(P, B, S);
    }
    if (t == 0) {
      CUL(ensure(L.rescue_list, (size_t)n_slots * 4)); CUL(ensure(L.verify_list, (size_t)n_slots * 8));  // verify_list: cluster's, then verify's
      CUL(cudaMemsetAsync(L.d_count + 1, 0, 3 * sizeof(int), st));
      CUL(cudaEventRecord(L.ev_sub[0], st));
      CUL(cudaEventRecord(L.ev_sub[1], st));
      {
        // first-pass tile: enough rows for a typical read (about 2L/(w+1) minimizers, most of them single hits)
        int rows0 = (2 * ctx->params.max_read_length / (ctx->w + 1) + 15) / 16 * 16;  // 16 rows at 2x50, 48 at 2x150
        rows0 = std::max(16, std::min(rows0, S.caps.hc));
        cluster_kernel<<<(2 * n_slots + CLUSTER_NT - 1) / CLUSTER_NT, CLUSTER_NT, (size_t)rows0 * CLUSTER_NT * 8, st>>>(P, ix, S, L.ctr, 0, rows0, (int *)L.verify_list.p, L.d_count + 3);
        cluster_passes = rows0 < S.caps.hc ? 2 : 1;
        if (rows0 < S.caps.hc)
          cluster_kernel<<<(2 * n_slots + CLUSTER_NT - 1) / CLUSTER_NT, CLUSTER_NT, (size_t)S.caps.hc * CLUSTER_NT * 8, st>>>(P, ix, S, L.ctr, 1, S.caps.hc, (int *)L.verify_list.p, L.d_count + 3);
      }
      CUL(cudaEventRecord(e1, st));

      pair_candidates_kernel<<<(n_slots + TB - 1) / TB, TB, 0, st>>>(P, ix, S, L.ctr, 0, (int *)L.rescue_list.p, L.d_count + 1);
      pair_candidates_kernel<<<(n_slots + 63) / 64, 64, 0, st>>>(P, ix, S, L.ctr, 1, (int *)L.rescue_list.p, L.d_count + 1);
      CUL(cudaEventRecord(e2, st));
      if (P.split) verify_split_kernel<<<(2 * n_slots + TB - 1) / TB, TB, 0, st>>>(P, R, B, S, L.ctr);
      else {
        verify_kernel<<<(2 * n_slots + TB - 1) / TB, TB, 0, st>>>(P, R, B, S, L.ctr, 0, (int *)L.verify_list.p, L.d_count + 2);
        verify_kernel<<<(2 * n_slots + 63) / 64, 64, (size_t)2 * S.caps.maxmm * 64, st>>>(P, R, B, S, L.ctr, 1, (int *)L.verify_list.p, L.d_count + 2);
      }
      CUL(cudaEventRecord(e3, st));
      if (P.split) pairing_split_kernel<<<(n_slots + TB - 1) / TB, TB, 0, st>>>(P, S, (int *)L.nbest.p);
      else pairing_kernel<<<(n_slots + TB - 1) / TB, TB, 0, st>>>(P, S, (int *)L.nbest.p);
      CUL(cudaEventRecord(e4, st));
    } else {  // overflow tiers: one CTA per read / pair; shared-memory sort buffers sized to the tier
      auto cap_of = [](int n) { int c = 1; while (c < n) c <<= 1; return std::min(c, CTA_SORT_SMEM_MAX); };
      const int c_seed = cap_of(2 * tier.caps.hc), c_pc = cap_of(tier.caps.hc), c_ver = cap_of(tier.caps.cc), c_pair = cap_of(tier.caps.mc);
      seed_cta_kernel<<<2 * n_slots, CTA_NT, (size_t)c_seed * 11 + (size_t)(tier.caps.maxmm + 1) * 12 + 16, st>>>(P, ix, B, S, L.tiers[0].view, L.ctr, c_seed);
      CUL(cudaEventRecord(e1, st));
      {
        const int lcap = std::min(tier.caps.cc, 512), fcap = 2 * tier.caps.cc;
        // the last tier's shared-memory lists leave room for two CTAs per SM only: sixteen warps each instead of four keep as
        // many dependent occurrence-list searches in flight as the smaller tiers do
        const size_t pc_smem = pair_candidates_cta_smem(c_pc, lcap, tier.caps.maxmm, fcap);
        const int pc_nt = pc_smem > 64 * 1024 ? PC_CTA_NT_MAX : CTA_NT;
        pair_candidates_cta_kernel<<<n_slots, pc_nt, pc_smem, st>>>(P, ix, S, L.ctr, c_pc, lcap, fcap, nullptr, nullptr);
      }
      CUL(cudaEventRecord(e2, st));
      if (P.split) verify_split_cta_kernel<<<2 * n_slots, CTA_NT, (size_t)c_ver * 9, st>>>(P, R, B, S, L.ctr, c_ver);
      else verify_cta_kernel<<<2 * n_slots, tier.caps.cc > 1024 ? VERIFY_NT_MAX : CTA_NT, (size_t)c_ver * 9 + 2 * (size_t)tier.caps.maxmm + 16, st>>>(P, R, B, S, L.ctr, c_ver);
      CUL(cudaEventRecord(e3, st));
      if (P.split) pairing_split_kernel<<<(n_slots + TB - 1) / TB, TB, 0, st>>>(P, S, (int *)L.nbest.p);
      else pairing_cta_kernel<<<n_slots, CTA_NT, (size_t)c_pair * 10, st>>>(P, S, (int *)L.nbest.p, c_pair);
      CUL(cudaEventRecord(e4, st));
    }
    // kernels launched above: tier 0 = front end (counted per piece) + cluster (1 or 2 passes) + pair_candidates x2 +
    // verify (x2 unless split) + pairing; overflow tiers = prep + four CTA kernels
    if (t == 0) acc.launches += cluster_passes + 2 + (P.split ? 1 : 2) + 1;
    else acc.launches += 5;
    CUL(ensure(tier.ovf_list, (size_t)n_slots * 4));
    CUL(cudaMemsetAsync(L.d_count, 0, sizeof(int), st));
    collect_overflow_kernel<<<(n_slots + 255) / 256, 256, 0, st>>>(S, (int *)tier.ovf_list.p, L.d_count);
    acc.launches += 1;
    int n_ovf = 0;
    CUL(cudaMemcpyAsync(&n_ovf, L.d_count, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUL(cudaStreamSynchronize(st));
    CUL(cudaGetLastError());
    float f;
    cudaEventElapsedTime(&f, e0, e1); acc.ms_seed += f;
    if (t == 0) { cudaEventElapsedTime(&f, e0, L.ev_sub[0]); acc.ms_minimizer += f; cudaEventElapsedTime(&f, L.ev_sub[0], L.ev_sub[1]); acc.ms_probe += f; cudaEventElapsedTime(&f, L.ev_sub[1], e1); acc.ms_cluster += f; }
    cudaEventElapsedTime(&f, e1, e2); acc.ms_pc += f;
    cudaEventElapsedTime(&f, e2, e3); acc.ms_ver += f;
    cudaEventElapsedTime(&f, e3, e4); acc.ms_pair += f;
    tiers_used = t + 1;
    if (n_ovf > 0 && t + 1 < N_TIERS) {
      // deterministic order for the next tier: sort the pair list (atomic append order is arbitrary)
      std::vector<int> h(n_ovf);
      CUL(cudaMemcpyAsync(h.data(), tier.ovf_list.p, (size_t)n_ovf * 4, cudaMemcpyDeviceToHost, st));
      CUL(cudaStreamSynchronize(st));
      std::sort(h.begin(), h.end());
      CUL(cudaMemcpyAsync(tier.ovf_list.p, h.data(), (size_t)n_ovf * 4, cudaMemcpyHostToDevice, st));
      CUL(cudaStreamSynchronize(st));
      pair_list = (const int *)tier.ovf_list.p;
    } else if (n_ovf > 0) {
      n_overflow_final = n_ovf;
    }
    n_slots = n_ovf;
  }
  // multi-mapper sampling: one warp per taskloop chunk of every reference batch in this lane
  std::vector<int> chunks;
  for (u32 b0 = 0; b0 < n; b0 += (u32)ctx->params.batch_size) taskloop_chunks(b0, std::min<u32>((u32)ctx->params.batch_size, n - b0), chunks);
  const int n_chunks = (int)chunks.size();
  chunks.push_back((int)n);
  CUL(ensure(L.chunk_start, chunks.size() * 4));
  CUL(cudaEventRecord(L.ev[2], st));
  CUL(cudaMemcpyAsync(L.chunk_start.p, chunks.data(), chunks.size() * 4, cudaMemcpyHostToDevice, st));
  select_kernel<<<(n_chunks + 3) / 4, 128, 0, st>>>(P, n_chunks, (const int *)L.chunk_start.p, (const int *)L.nbest.p, (int *)L.sel.p, ctx->mt_init);
  CUL(cudaEventRecord(L.ev[3], st));
  // the overflow tiers' emits (few pairs, long per-thread sweeps) run beside tier 0's on their own streams
  if (tiers_used > 1) CUL(cudaEventRecord(L.ev_fork, st));
  for (int t = tiers_used - 1; t >= 0; --t) {
    const Scratch S = L.tiers[t].view;
    cudaStream_t es = t == 0 ? st : L.aux[t - 1];
    if (t > 0) CUL(cudaStreamWaitEvent(es, L.ev_fork, 0));
    if (P.split) emit_split_kernel<<<(S.n_slots + TB - 1) / TB, TB, 0, es>>>(P, R, B, T, S, (const int *)L.sel.p, (OutPairs *)L.out_rec.p, (int *)L.out_n.p, L.ctr);
    else if (sam && P.se) emit_sam_se_kernel<<<(S.n_slots + TB - 1) / TB, TB, 0, es>>>(P, R, B, T, S, (const int *)L.sel.p, (OutSam *)L.out_rec.p, (int *)L.out_n.p, L.ctr);
    else if (sam) emit_sam_kernel<<<(S.n_slots + TB - 1) / TB, TB, 0, es>>>(P, R, B, T, S, (const int *)L.sel.p, (OutSam *)L.out_rec.p, (int *)L.out_n.p, L.ctr);
    else if (P.se) emit_se_kernel<<<(S.n_slots + TB - 1) / TB, TB, 0, es>>>(P, R, B, T, S, (const int *)L.sel.p, (OutRecord *)L.out_rec.p, (int *)L.out_n.p, L.ctr);
    else if (t > 0) emit_cta_kernel<<<S.n_slots, CTA_NT, 0, es>>>(P, R, B, T, S, (const int *)L.sel.p, (OutRecord *)L.out_rec.p, (int *)L.out_n.p, L.ctr);
    else {
      CUL(ensure(L.emit_list, (size_t)S.n_slots * mb * sizeof(int4)));
      CUL(cudaMemsetAsync(L.d_count + 2, 0, sizeof(int), es));  // (verify's list counter: free by now)
      emit_kernel<<<(S.n_slots + TB - 1) / TB, TB, 0, es>>>(P, R, B, T, S, (const int *)L.sel.p, (OutRecord *)L.out_rec.p, (int *)L.out_n.p, L.ctr, (int4 *)L.emit_list.p,
                                                           L.d_count + 2);
      // the pairs whose start coordinates need the bit-vector traceback (indels); the grid covers the worst case, idle threads leave at once
      emit_dp_kernel<<<(unsigned)(((size_t)S.n_slots * mb + TB - 1) / TB), TB, 0, es>>>(P, R, B, T, S, (OutRecord *)L.out_rec.p, (const int4 *)L.emit_list.p, L.d_count + 2);
      acc.launches += 1;
    }
    if (t > 0) CUL(cudaEventRecord(L.ev_join[t - 1], es));
  }
  for (int t = 1; t < tiers_used; ++t) CUL(cudaStreamWaitEvent(st, L.ev_join[t - 1], 0));
  acc.launches += 1 + tiers_used;
  // read-order compaction
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, (const int *)L.out_n.p, (u64 *)L.offs.p, (int)n + 1, st);
  CUL(ensure(L.cub_tmp, tmp_bytes));
  // out_n has n entries plus one trailing zero so that offs[n] = total
  cub::DeviceScan::ExclusiveSum(L.cub_tmp.p, tmp_bytes, (const int *)L.out_n.p, (u64 *)L.offs.p, (int)n + 1, st);
  if (sam) compact_words_kernel<<<(n + 255) / 256, 256, 0, st>>>((int)n, mb, (int)(rec_bytes / 4), (const u32 *)L.out_rec.p, (const int *)L.out_n.p, (const u64 *)L.offs.p, (u32 *)dst);
  else compact_kernel<<<(n + 255) / 256, 256, 0, st>>>((int)n, mb, (const OutRecord *)L.out_rec.p, (const int *)L.out_n.p, (const u64 *)L.offs.p, dst);
  acc.launches += 2;
  if (bc_dst) {
    compact_bc_kernel<<<(n + 255) / 256, 256, 0, st>>>((int)n, (const int *)L.out_n.p, (const u64 *)L.offs.p, (const u64 *)L.bc_key.p, bc_dst);
    acc.launches += 1;
  }
  u64 total = 0;
  CUL(cudaEventRecord(L.ev[4], st));
  CUL(cudaMemcpyAsync(&total, (u64 *)L.offs.p + n, 8, cudaMemcpyDeviceToHost, st));
  Counters hc;
  CUL(cudaMemcpyAsync(&hc, L.ctr, sizeof(Counters), cudaMemcpyDeviceToHost, st));
  CUL(cudaStreamSynchronize(st));
  CUL(cudaGetLastError());
  float f;
  cudaEventElapsedTime(&f, L.ev[2], L.ev[3]); acc.ms_select += f;
  cudaEventElapsedTime(&f, L.ev[3], L.ev[4]); acc.ms_emit += f;
  {
    u64 *a = (u64 *)&acc.c;
    const u64 *b = (const u64 *)&hc;
    for (size_t i = 0; i < sizeof(Counters) / 8; ++i) a[i] += b[i];
  }
  for (int t = 0; t < N_TIERS; ++t) if (t < tiers_used) acc.tier_pairs[t] += (u64)L.tiers[t].n_slots;
  acc.tiers_used = std::max(acc.tiers_used, tiers_used);
  acc.n_overflow += n_overflow_final + (sam ? hc.n_overflow : 0);  // SAM: reads / CIGARs beyond the fixed record
  J.total = total;
  L.tiers_used = tiers_used;
  return CMX_OK;
}

int cmx_map_batch_pe(cmx_ctx *ctx, const cmx_batch *in, cmx_records *out, void *user_stream) {
  if (!ctx || !in || !out) return CMX_ERR_INVALID;
  if (!ctx->slots || !ctx->ref_seq) return fail(ctx, CMX_ERR_STATE, "index and reference must be uploaded first");
  const u32 n = in->n_pairs;
  const int mb = ctx->params.max_num_best_mappings;
  out->n_records = 0; out->n_mapped_pairs = out->n_uniquely_mapped_pairs = out->n_candidates = out->n_overflow_pairs = 0;
  out->n_barcodes_in_whitelist = out->n_barcodes_corrected = 0;
  if (in->bc_seq && (in->bc_len == 0 || in->bc_len > 32 || !in->bc_qual)) return fail(ctx, CMX_ERR_INVALID, "barcodes need bc_qual and 1 <= bc_len <= 32");
  if (in->bc_seq && ctx->params.split_alignment) return fail(ctx, CMX_ERR_INVALID, "barcodes are not supported with split alignment");
  const bool se = ctx->params.single_end != 0;
  if (!in->seq1 || !in->off1 || (!se && (!in->seq2 || !in->off2))) return fail(ctx, CMX_ERR_INVALID, "cmx_batch: read pointers missing");
  if (n == 0) return CMX_OK;
  if (out->capacity < (u64)n * mb) return fail(ctx, CMX_ERR_INVALID, "records capacity %llu < n_pairs*max_num_best_mappings", (unsigned long long)out->capacity);
  CU(cudaSetDevice(ctx->device));
  (void)user_stream;  // the context's own streams are used; the call is synchronous
  cudaStream_t up = ctx->up_stream;
  const u32 bs = (u32)ctx->params.batch_size;
  const u32 n_sub = (n + bs - 1) / bs;  // reference batches in this call
  const bool bc = in->bc_seq && in->bc_len;
  CU(cudaEventRecord(ctx->ev[0], up));
  // ---- inputs: device pointers as they are; host buffers go up in pieces (one per reference batch) on the upload
  // stream, so the first kernels start on piece 0 while the other pieces are still in flight
  DevBatch B{};
  B.n_pairs = n; B.first_read_id = in->first_read_id;
  const u8 *bcs = nullptr, *bcq = nullptr;
  const bool pieces = !in->on_device;
  // upload granularity: a quarter of a reference batch, so the first kernels start after a sixteenth of a 4-batch upload
  const u32 ps = (bs % 4 == 0 && bs / 4 >= 32768) ? bs / 4 : bs;
  u32 n_pieces = 0;
  if (in->on_device) {
    B.seq1 = (const u8 *)in->seq1; B.off1 = in->off1; B.seq2 = (const u8 *)in->seq2; B.off2 = in->off2;
    if (bc) { bcs = (const u8 *)in->bc_seq; bcq = (const u8 *)in->bc_qual; }
  } else {
    const size_t b1 = in->off1[n], b2 = se ? 0 : in->off2[n];
    CU(ensure(ctx->seq1, b1 + 64)); CU(ensure(ctx->seq2, b2 + 64));
    CU(ensure(ctx->off1, (size_t)(n + 1) * 4)); CU(ensure(ctx->off2, (size_t)(n + 1) * 4));
    n_pieces = (n + ps - 1) / ps;
    while (ctx->ev_up.size() < n_pieces) { cudaEvent_t e; CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); ctx->ev_up.push_back(e); }
    CU(cudaMemcpyAsync(ctx->off1.p, in->off1, (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, up));
    if (!se) CU(cudaMemcpyAsync(ctx->off2.p, in->off2, (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, up));
    if (bc) {
      CU(ensure(ctx->bc_seq, (size_t)n * in->bc_len + 16)); CU(ensure(ctx->bc_qual, (size_t)n * in->bc_len + 16));
      CU(cudaMemcpyAsync(ctx->bc_seq.p, in->bc_seq, (size_t)n * in->bc_len, cudaMemcpyHostToDevice, up));
      CU(cudaMemcpyAsync(ctx->bc_qual.p, in->bc_qual, (size_t)n * in->bc_len, cudaMemcpyHostToDevice, up));
      if (!ctx->ev_bc) CU(cudaEventCreateWithFlags(&ctx->ev_bc, cudaEventDisableTiming));
      CU(cudaEventRecord(ctx->ev_bc, up));
      bcs = (const u8 *)ctx->bc_seq.p; bcq = (const u8 *)ctx->bc_qual.p;
    }
    for (u32 s = 0; s < n_pieces; ++s) {
      const u32 p0 = s * ps, p1 = std::min(n, p0 + ps);
      CU(cudaMemcpyAsync((char *)ctx->seq1.p + in->off1[p0], in->seq1 + in->off1[p0], in->off1[p1] - in->off1[p0], cudaMemcpyHostToDevice, up));
      if (!se) CU(cudaMemcpyAsync((char *)ctx->seq2.p + in->off2[p0], in->seq2 + in->off2[p0], in->off2[p1] - in->off2[p0], cudaMemcpyHostToDevice, up));
      CU(cudaEventRecord(ctx->ev_up[s], up));
    }
    B.seq1 = (const u8 *)ctx->seq1.p; B.off1 = (const u32 *)ctx->off1.p;
    B.seq2 = (const u8 *)ctx->seq2.p; B.off2 = (const u32 *)ctx->off2.p;
  }
  CU(cudaEventRecord(ctx->ev[1], up));
  // ---- lanes: contiguous groups of whole reference batches (the sampling generator restarts per batch chunk,
  // so a lane needs nothing from its neighbours)
  const int n_lanes = (int)std::min<u32>((u32)ctx->n_lanes, n_sub);
  LaneJob jobs[CMX_MAX_LANES];
  for (int l = 0; l < n_lanes; ++l) {
    const u32 s0 = (u32)((u64)n_sub * l / n_lanes), s1 = (u32)((u64)n_sub * (l + 1) / n_lanes);
    LaneJob &J = jobs[l];
    J.p0 = s0 * bs; J.n = std::min(n, s1 * bs) - J.p0;
    if (pieces) { J.piece = ps; J.piece0 = s0 * (bs / ps); J.piece_ready = &ctx->ev_up; }
    if (bc) { J.bc_seq = bcs + (size_t)J.p0 * in->bc_len; J.bc_qual = bcq + (size_t)J.p0 * in->bc_len; J.bc_len = in->bc_len; J.bc_ready = pieces ? ctx->ev_bc : nullptr; }
    J.want_bc = bc && out->barcode_keys;
    if (out->on_device && n_lanes == 1) J.dst = (OutRecord *)out->records;
  }
  // every lane maps its pairs and then delivers its records itself, behind the records of the lanes before it (their
  // counts are known as soon as those lanes have finished mapping), so the copies of early lanes overlap later lanes
  std::atomic<int> mapped[CMX_MAX_LANES];
  for (auto &f : mapped) f.store(0);
  const size_t rec_bytes = ctx->params.output_format == 4 ? sizeof(OutSam) : sizeof(OutRecord);
  auto lane_main = [&](int l) {
    LaneJob &J = jobs[l];
    Lane &L = ctx->lanes[l];
    J.rc = run_lane(ctx, L, B, J);
    if (J.rc) J.total = 0;
    mapped[l].store(1, std::memory_order_release);
    u64 before = 0;
    for (int k = 0; k < l; ++k) {
      while (!mapped[k].load(std::memory_order_acquire)) std::this_thread::yield();
      before += jobs[k].total;
    }
    if (J.rc || !J.total) return;
    cudaError_t e = cudaSuccess;
    if ((void *)J.dst != (void *)out->records)
      e = cudaMemcpyAsync((char *)out->records + before * rec_bytes, J.dst, J.total * rec_bytes, out->on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, L.stream);
    if (e == cudaSuccess && J.want_bc) e = cudaMemcpyAsync(out->barcode_keys + before, L.bc_out.p, J.total * 8, cudaMemcpyDeviceToHost, L.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(L.stream);
    if (e != cudaSuccess) { L.err = std::string("record copy: ") + cudaGetErrorString(e); J.rc = CMX_ERR_CUDA; }
  };
  CU(cudaEventRecord(ctx->ev[2], ctx->down_stream));
  {
    std::vector<std::thread> th;
    for (int l = 1; l < n_lanes; ++l) th.emplace_back(lane_main, l);
    lane_main(0);
    for (auto &t : th) t.join();
  }
  CU(cudaEventRecord(ctx->ev[3], ctx->down_stream));
  CU(cudaStreamSynchronize(ctx->down_stream));
  ctx->last_lanes_used = n_lanes;
  ctx->last_n_pairs = n;
  for (int l = 0; l < n_lanes; ++l)
    if (jobs[l].rc) { ctx->err = ctx->lanes[l].err; return jobs[l].rc; }
  u64 total = 0;
  for (int l = 0; l < n_lanes; ++l) total += jobs[l].total;
  CU(cudaGetLastError());
  float ms_h2d = 0, ms_call = 0;
  cudaEventElapsedTime(&ms_h2d, ctx->ev[0], ctx->ev[1]);
  cudaEventElapsedTime(&ms_call, ctx->ev[2], ctx->ev[3]);  // lanes launched .. last record delivered (events on an idle stream)
  BatchAcc acc;
  for (int l = 0; l < n_lanes; ++l) {
    const BatchAcc &a = jobs[l].acc;
    acc.ms_seed += a.ms_seed; acc.ms_pc += a.ms_pc; acc.ms_ver += a.ms_ver; acc.ms_pair += a.ms_pair; acc.ms_minimizer += a.ms_minimizer;
    acc.ms_probe += a.ms_probe; acc.ms_cluster += a.ms_cluster; acc.ms_select += a.ms_select; acc.ms_emit += a.ms_emit;
    acc.launches += a.launches; acc.n_overflow += a.n_overflow;
    u64 *x = (u64 *)&acc.c;
    const u64 *y = (const u64 *)&a.c;
    for (size_t i = 0; i < sizeof(Counters) / 8; ++i) x[i] += y[i];
    for (int t = 0; t < N_TIERS; ++t) acc.tier_pairs[t] += a.tier_pairs[t];
  }
  out->n_records = total;
  out->n_mapped_pairs = acc.c.n_mapped; out->n_uniquely_mapped_pairs = acc.c.n_unique; out->n_candidates = acc.c.n_candidates;
  out->n_overflow_pairs = acc.n_overflow;
  out->n_barcodes_in_whitelist = acc.c.n_bc_in_whitelist; out->n_barcodes_corrected = acc.c.n_bc_corrected;
  cmx_timing &tm = ctx->timing;
  memset(&tm, 0, sizeof(tm));
  // stage times are sums over the lanes' own streams; lanes overlap, so they add up to more than total_ms
  tm.h2d_ms = ms_h2d; tm.d2h_ms = 0;  // record copies run on the lanes' own streams, overlapped with other lanes' kernels
  tm.seed_ms = acc.ms_seed; tm.front_ms = acc.ms_minimizer + acc.ms_probe; tm.reserved_ms = 0; tm.cluster_ms = acc.ms_cluster;
  tm.pair_candidates_ms = acc.ms_pc; tm.verify_ms = acc.ms_ver; tm.pairing_ms = acc.ms_pair; tm.select_ms = acc.ms_select; tm.emit_ms = acc.ms_emit;
  tm.total_ms = ms_call;
  tm.n_minimizers = acc.c.n_minimizers; tm.n_probe_steps = acc.c.n_probe_steps; tm.n_found = acc.c.n_found; tm.n_occ_reads = acc.c.n_occ_reads;
  tm.n_verified = acc.c.n_verified; tm.n_launches = acc.launches;
  for (int t = 0; t < 3; ++t) tm.tier_pairs[t] = acc.tier_pairs[t];
  for (int r = 0; r < 8; ++r) tm.escalations[r] = acc.c.ovf_reason[r];
  if (acc.n_overflow) return fail(ctx, CMX_ERR_OVERFLOW, "%llu pair(s) exceeded the largest scratch tier", (unsigned long long)acc.n_overflow);
  return CMX_OK;
}

// Host helper: bytes taken by the first min(max_records, complete) 4-line records of `text` (a record is complete when
// its fourth newline is present).
uint64_t cmx_fastq_cut(const char *text, uint64_t n_bytes, uint32_t max_records, uint32_t *n_records) {
  uint64_t pos = 0, end_of_last = 0;
  uint32_t lines = 0, recs = 0;
  while (recs < max_records) {
    const void *q = memchr(text + pos, '\n', n_bytes - pos);
    if (!q) break;
    pos = (uint64_t)((const char *)q - text) + 1;
    if (++lines == 4) { lines = 0; ++recs; end_of_last = pos; }
  }
  if (n_records) *n_records = recs;
  return end_of_last;
}

int cmx_ingest_fastq(cmx_ctx *ctx, int slot, const char *text, uint64_t n_bytes, int want_qual, uint32_t *name_spans, cmx_ingested *out) {
  if (!ctx || !out || slot < 0 || slot >= CMX_INGEST_SLOTS || (!text && n_bytes)) return CMX_ERR_INVALID;
  memset(out, 0, sizeof(*out));
  if (n_bytes == 0) return CMX_OK;
  if (n_bytes >= 0xFFFFFFF0ull) return fail(ctx, CMX_ERR_INVALID, "cmx_ingest_fastq: chunk of 4 GiB or more");
  if (text[n_bytes - 1] != '\n') return fail(ctx, CMX_ERR_INVALID, "cmx_ingest_fastq: the chunk must end at a record boundary (cmx_fastq_cut)");
  CU(cudaSetDevice(ctx->device));
  IngestSlot &g = ctx->ingest[slot];
  if (!g.stream) CU(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
  cudaStream_t st = g.stream;
  const u32 nb = (u32)n_bytes;
  CU(ensure(g.text, n_bytes + 16)); CU(ensure(g.nl, ((size_t)nb + 16) * 4)); CU(ensure(g.count, 16)); CU(ensure(g.stats, sizeof(IngestStats)));
  CU(cudaMemcpyAsync(g.text.p, text, n_bytes, cudaMemcpyHostToDevice, st));
  // newline positions: select the indices whose byte is '\n'
  thrust::counting_iterator<u32> idx(0);
  CU(ensure(g.qual_start, (size_t)nb));  // newline flags live here until the per-record arrays are sized (same buffer, reused)
  u8 *flags = (u8 *)g.qual_start.p;
  newline_flag_kernel<<<(nb + 255) / 256, 256, 0, st>>>((const char *)g.text.p, nb, flags);
  size_t tb = 0;
  cub::DeviceSelect::Flagged(nullptr, tb, idx, flags, (u32 *)g.nl.p, (u32 *)g.count.p, (int)nb, st);
  CU(ensure(g.tmp, tb));
  CU(cub::DeviceSelect::Flagged(g.tmp.p, tb, idx, flags, (u32 *)g.nl.p, (u32 *)g.count.p, (int)nb, st));
  u32 n_nl = 0;
  CU(cudaMemcpyAsync(&n_nl, g.count.p, 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (n_nl % 4 != 0) return fail(ctx, CMX_ERR_INVALID, "cmx_ingest_fastq: %u lines is not a whole number of 4-line records", n_nl);
  const u32 n = n_nl / 4;
  CU(ensure(g.seq_start, (size_t)n * 4)); CU(ensure(g.qual_start, (size_t)n * 4)); CU(ensure(g.len, (size_t)(n + 1) * 4)); CU(ensure(g.off, (size_t)(n + 1) * 4));
  if (name_spans) CU(ensure(g.spans, (size_t)n * 8));
  IngestStats hs = {0, 0, 0, 0, 0xFFFFFFFFu, 0};
  CU(cudaMemcpyAsync(g.stats.p, &hs, sizeof(hs), cudaMemcpyHostToDevice, st));
  CU(cudaMemsetAsync((u32 *)g.len.p + n, 0, 4, st));
  ingest_record_kernel<<<(n + 255) / 256, 256, 0, st>>>((const char *)g.text.p, (const u32 *)g.nl.p, n, (u32 *)g.seq_start.p, (u32 *)g.qual_start.p, (u32 *)g.len.p,
                                                        name_spans ? (u32 *)g.spans.p : nullptr, (IngestStats *)g.stats.p);
  cub::DeviceScan::ExclusiveSum(nullptr, tb, (const u32 *)g.len.p, (u32 *)g.off.p, (int)n + 1, st);
  CU(ensure(g.tmp, tb));
  CU(cub::DeviceScan::ExclusiveSum(g.tmp.p, tb, (const u32 *)g.len.p, (u32 *)g.off.p, (int)n + 1, st));
  // sequence bytes never exceed the chunk; a malformed record (quality shorter than sequence) is only reported after the
  // pack kernel has run, so the buffers are sized by the chunk and the quality copy is bounded by the quality line itself
  CU(ensure(g.seq, n_bytes + 64));
  if (want_qual) CU(ensure(g.qual, n_bytes + 64));
  ingest_pack_kernel<<<(unsigned)(((u64)n * 32 + 255) / 256), 256, 0, st>>>((const char *)g.text.p, (const u32 *)g.seq_start.p, (const u32 *)g.qual_start.p,
                                                                           (const u32 *)g.off.p, (const u32 *)g.nl.p, n, (char *)g.seq.p, want_qual ? (char *)g.qual.p : nullptr);
  CU(cudaMemcpyAsync(&hs, g.stats.p, sizeof(hs), cudaMemcpyDeviceToHost, st));
  if (name_spans) CU(cudaMemcpyAsync(name_spans, g.spans.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  CU(cudaGetLastError());
  if (hs.bad_header || hs.bad_plus) return fail(ctx, CMX_ERR_INVALID, "cmx_ingest_fastq: not 4-line FASTQ (%u headers without '@', %u separator lines without '+')", hs.bad_header, hs.bad_plus);
  if (hs.empty_reads) return fail(ctx, CMX_ERR_INVALID, "cmx_ingest_fastq: %u empty reads (the reference skips them per file; use the host reader)", hs.empty_reads);
  if (hs.qual_mismatch) return fail(ctx, CMX_ERR_INVALID, "cmx_ingest_fastq: %u records whose quality and sequence lengths differ", hs.qual_mismatch);
  out->n_reads = n; out->seq = (const char *)g.seq.p; out->off = (const uint32_t *)g.off.p; out->qual = want_qual ? (const char *)g.qual.p : nullptr;
  out->min_len = n ? hs.min_len : 0; out->max_len = hs.max_len;
  return CMX_OK;
}

int cmx_host_register(void *ptr, uint64_t bytes) {
  if (!ptr || !bytes) return CMX_ERR_INVALID;
  return cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable) == cudaSuccess ? CMX_OK : CMX_ERR_CUDA;
}
int cmx_host_unregister(void *ptr) {
  if (!ptr) return CMX_ERR_INVALID;
  return cudaHostUnregister(ptr) == cudaSuccess ? CMX_OK : CMX_ERR_CUDA;
}

int cmx_set_lanes(cmx_ctx *ctx, int n_lanes) {
  if (!ctx || n_lanes < 1 || n_lanes > CMX_MAX_LANES) return CMX_ERR_INVALID;
  ctx->n_lanes = n_lanes;
  return CMX_OK;
}

int cmx_last_batch_timing(cmx_ctx *ctx, cmx_timing *out) {
  if (!ctx || !out) return CMX_ERR_INVALID;
  *out = ctx->timing;
  return CMX_OK;
}

__global__ void trace_kernel(Scratch S, cmx_pair_trace *out, u32 p0) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= S.n_slots) return;
  const PairMeta &pm = S.pmeta[slot];
  if (pm.status == ST_OVERFLOW) return;
  cmx_pair_trace t;
  memset(&t, 0, sizeof(t));
  const ReadMeta *rm = S.rmeta + 2 * slot;
  for (int m = 0; m < 2; ++m) {
    t.trimmed_len[m] = rm[m].len;
    t.n_minimizers[m] = rm[m].n_mm;
    t.n_pos_candidates_gen[m] = rm[m].n_cand_gen[0]; t.n_neg_candidates_gen[m] = rm[m].n_cand_gen[1];
    t.n_pos_candidates[m] = rm[m].n_cand[0]; t.n_neg_candidates[m] = rm[m].n_cand[1];
    t.n_pos_mappings[m] = rm[m].n_map[0]; t.n_neg_mappings[m] = rm[m].n_map[1];
    t.min_errors[m] = rm[m].min_err; t.second_min_errors[m] = rm[m].second_min_err;
    t.n_best[m] = rm[m].n_best; t.n_second_best[m] = rm[m].n_second_best;
    t.repetitive_seed_length[m] = rm[m].rep_len;
  }
  t.supplement_result = pm.sup;
  t.min_sum_errors = pm.min_sum; t.second_min_sum_errors = pm.second_min_sum; t.n_best_pairs = pm.n_best; t.n_second_best_pairs = pm.n_second_best;
  t.n_records = pm.n_rec;
  out[p0 + slot_pair(S, slot)] = t;
}

int cmx_last_batch_trace(cmx_ctx *ctx, cmx_pair_trace *out, uint32_t n_pairs) {
  if (!ctx || !out || n_pairs != ctx->last_n_pairs) return CMX_ERR_INVALID;
  CU(cudaSetDevice(ctx->device));
  CU(ensure(ctx->trace, (size_t)n_pairs * sizeof(cmx_pair_trace)));
  CU(cudaMemset(ctx->trace.p, 0, (size_t)n_pairs * sizeof(cmx_pair_trace)));
  for (int l = 0; l < ctx->last_lanes_used; ++l) {
    const Lane &L = ctx->lanes[l];
    for (int t = 0; t < L.tiers_used; ++t) {
      const Scratch S = L.tiers[t].view;
      trace_kernel<<<(S.n_slots + 255) / 256, 256>>>(S, (cmx_pair_trace *)ctx->trace.p, L.p0);
    }
  }
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(out, ctx->trace.p, (size_t)n_pairs * sizeof(cmx_pair_trace), cudaMemcpyDeviceToHost));
  return CMX_OK;
}

// ---- stage entry points ---------------------------------------------------------------------------
__global__ void stage_minimizers_kernel(DevBatch B, int k, int w, u64 *out_hash, u32 *out_pos, int *out_n, u32 stride) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= 2 * (int)B.n_pairs) return;
  const int pair = r >> 1, mate = r & 1;
  const u8 *seq = read_ptr(B, pair, mate);
  u64 *oh = out_hash + (size_t)r * stride;
  u32 *op = out_pos + (size_t)r * stride;
  int n = 0;
  minimizer_scan_any([&](int i) { return seq[i]; }, read_raw_len(B, pair, mate), k, w, [&](u64 h, u32 p) { if (n < (int)stride) { oh[n] = h; op[n] = p; } ++n; });
  out_n[r] = n;
}

int cmx_stage_minimizers(cmx_ctx *ctx, const cmx_batch *in, uint64_t *out_hash, uint32_t *out_pos, int32_t *out_n, uint32_t stride) {
  if (!ctx || !in || !out_hash || !out_pos || !out_n) return CMX_ERR_INVALID;
  if (!ctx->k) return fail(ctx, CMX_ERR_STATE, "index (k, w) not set");
  CU(cudaSetDevice(ctx->device));
  DevBatch B{};
  int rc = upload_batch(ctx, in, &B);
  if (rc) return rc;
  const size_t R = 2 * (size_t)in->n_pairs;
  u64 *dh; u32 *dp; int *dn;
  CU(cudaMalloc(&dh, R * stride * 8)); CU(cudaMalloc(&dp, R * stride * 4)); CU(cudaMalloc(&dn, R * 4));
  stage_minimizers_kernel<<<(unsigned)((R + 127) / 128), 128, 0, ctx->stream>>>(B, ctx->k, ctx->w, dh, dp, dn, stride);
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaGetLastError());
  CU(cudaMemcpy(out_hash, dh, R * stride * 8, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(out_pos, dp, R * stride * 4, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(out_n, dn, R * 4, cudaMemcpyDeviceToHost));
  cudaFree(dh); cudaFree(dp); cudaFree(dn);
  return CMX_OK;
}

__global__ void stage_probe_kernel(DevIndex ix, const u64 *h, u64 n, u8 *found, u64 *key, u64 *val) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u64 v = 0;
  int steps;
  const int kind = index_lookup(ix, h[i], &v, &steps);
  found[i] = kind != 0;
  key[i] = kind ? ((h[i] << 1) | (kind == 1 ? 1u : 0u)) : 0;
  val[i] = v;
}

int cmx_stage_probe(cmx_ctx *ctx, const uint64_t *hashes, uint64_t n, uint8_t *found, uint64_t *key, uint64_t *val) {
  if (!ctx || !hashes || !found || !key || !val) return CMX_ERR_INVALID;
  if (!ctx->slots) return fail(ctx, CMX_ERR_STATE, "index not uploaded");
  CU(cudaSetDevice(ctx->device));
  DevIndex ix;
  ix.slots = ctx->slots; ix.n_slots_mask = ctx->n_slots - 1; ix.shift = table_shift(ctx->n_slots); ix.occ = ctx->occ; ix.n_occ = ctx->n_occ; ix.k = ctx->k; ix.w = ctx->w;
  u64 *dh, *dk, *dv; u8 *df;
  CU(cudaMalloc(&dh, n * 8 + 8)); CU(cudaMalloc(&dk, n * 8 + 8)); CU(cudaMalloc(&dv, n * 8 + 8)); CU(cudaMalloc(&df, n + 8));
  CU(cudaMemcpy(dh, hashes, n * 8, cudaMemcpyHostToDevice));
  if (n) stage_probe_kernel<<<(unsigned)((n + 255) / 256), 256>>>(ix, dh, n, df, dk, dv);
  CU(cudaDeviceSynchronize());
  CU(cudaGetLastError());
  CU(cudaMemcpy(found, df, n, cudaMemcpyDeviceToHost)); CU(cudaMemcpy(key, dk, n * 8, cudaMemcpyDeviceToHost)); CU(cudaMemcpy(val, dv, n * 8, cudaMemcpyDeviceToHost));
  cudaFree(dh); cudaFree(dk); cudaFree(dv); cudaFree(df);
  return CMX_OK;
}

__global__ void stage_align_kernel(int e, int L, const u8 *pat, const u8 *txt, u64 n, int *err, int *endp) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u8 *p = pat + i * (size_t)(L + 2 * e), *t = txt + i * (size_t)L;
  int ep = 0;
  err[i] = banded_align(e, L, [&](int j) { return base_code(p[j]); }, [&](int j) { return base_code(t[j]); }, &ep);
  endp[i] = ep;
}

int cmx_stage_banded_align(cmx_ctx *ctx, int e, int read_len, const char *patterns, const char *texts, uint64_t n, int32_t *num_errors, int32_t *end_pos) {
  if (!ctx || !patterns || !texts || !num_errors || !end_pos || e < 1 || e > 15 || read_len < 1) return CMX_ERR_INVALID;
  CU(cudaSetDevice(ctx->device));
  u8 *dp, *dt; int *de, *dq;
  const size_t pb = n * (size_t)(read_len + 2 * e), tb = n * (size_t)read_len;
  CU(cudaMalloc(&dp, pb + 8)); CU(cudaMalloc(&dt, tb + 8)); CU(cudaMalloc(&de, n * 4 + 8)); CU(cudaMalloc(&dq, n * 4 + 8));
  CU(cudaMemcpy(dp, patterns, pb, cudaMemcpyHostToDevice)); CU(cudaMemcpy(dt, texts, tb, cudaMemcpyHostToDevice));
  if (n) stage_align_kernel<<<(unsigned)((n + 127) / 128), 128>>>(e, read_len, dp, dt, n, de, dq);
  CU(cudaDeviceSynchronize());
  CU(cudaGetLastError());
  CU(cudaMemcpy(num_errors, de, n * 4, cudaMemcpyDeviceToHost)); CU(cudaMemcpy(end_pos, dq, n * 4, cudaMemcpyDeviceToHost));
  cudaFree(dp); cudaFree(dt); cudaFree(de); cudaFree(dq);
  return CMX_OK;
}

__global__ void __launch_bounds__(CTA_NT) stage_cta_sort_kernel(u64 *keys, u8 *tags, int n, int sm_cap, int with_tags) {
  extern __shared__ u64 smk_stage[];
  u8 *smt = (u8 *)(smk_stage + sm_cap);
  if (with_tags) {
    auto cless = [](u64 pa, u8 ca, u64 pb, u8 cb) { return ca != cb ? ca > cb : pa < pb; };  // candidate order
    cta_sort_pairs<u8>(keys, tags, n, ~0ull, (u8)0, cless, smk_stage, smt, sm_cap);
  } else {
    cta_sort_keys(keys, n, smk_stage, sm_cap);
  }
}

int cmx_stage_cta_sort(cmx_ctx *ctx, uint64_t *keys, uint8_t *tags, uint32_t n, uint32_t sm_cap) {
  if (!ctx || !keys || sm_cap < 2 || (sm_cap & (sm_cap - 1)) || sm_cap > CTA_SORT_SMEM_MAX) return CMX_ERR_INVALID;
  CU(cudaSetDevice(ctx->device));
  size_t cap = 1;
  while (cap < n) cap <<= 1;
  u64 *dk; u8 *dt;
  CU(cudaMalloc(&dk, cap * 8 + 8)); CU(cudaMalloc(&dt, cap + 8));
  CU(cudaMemcpy(dk, keys, (size_t)n * 8, cudaMemcpyHostToDevice));
  if (tags) CU(cudaMemcpy(dt, tags, n, cudaMemcpyHostToDevice));
  stage_cta_sort_kernel<<<1, CTA_NT, (size_t)sm_cap * 9>>>(dk, dt, (int)n, (int)sm_cap, tags ? 1 : 0);
  CU(cudaDeviceSynchronize());
  CU(cudaGetLastError());
  CU(cudaMemcpy(keys, dk, (size_t)n * 8, cudaMemcpyDeviceToHost));
  if (tags) CU(cudaMemcpy(tags, dt, n, cudaMemcpyDeviceToHost));
  cudaFree(dk); cudaFree(dt);
  return CMX_OK;
}

// ---- post-processing (host side of the writer; GPU sort is a later row) ----------------------------
static inline auto rec_key(const cmx_pe_record &r) {  // bed_mapping.h:208-215 prefixed by rid
  return std::make_tuple(r.rid, r.fragment_start, r.fragment_length, r.mapq, r.direction, r.is_unique, r.read_id,
                         r.positive_alignment_length, r.negative_alignment_length);
}
static inline void tn5(cmx_pe_record &r) {  // bed_mapping.h:225-230
  r.fragment_start += 4; r.positive_alignment_length -= 4; r.fragment_length -= 9; r.negative_alignment_length -= 5;
}
static inline void tn5_se(cmx_pe_record &r) {  // single-end, bed_mapping.h:97-103
  if (r.direction == 1) r.fragment_start += 4; else r.fragment_length -= 5;
}

int cmx_postprocess(cmx_ctx *ctx, cmx_pe_record *recs, uint64_t n, uint64_t *n_out) {
  if (!ctx || (!recs && n) || !n_out) return CMX_ERR_INVALID;
  const cmx_params &p = ctx->params;
  *n_out = 0;
  if (n == 0) return CMX_OK;
  auto less = [](const cmx_pe_record &a, const cmx_pe_record &b) { return rec_key(a) < rec_key(b); };
  const bool se = p.single_end != 0;  // single-end duplicates: same start on the same sequence (bed_mapping.h:89-92)
  auto same = [se](const cmx_pe_record &a, const cmx_pe_record &b) {
    return a.rid == b.rid && a.fragment_start == b.fragment_start && (se || a.fragment_length == b.fragment_length);
  };
  auto tn5 = [se](cmx_pe_record &r) { if (se) tn5_se(r); else ::tn5(r); };
  uint64_t o = 0;
  if (p.low_memory_mode) {  // mapping_writer.h:166-376
    std::sort(recs, recs + n, less);
    uint64_t i = 0;
    while (i < n) {
      cmx_pe_record keep = recs[i];
      uint32_t dups = 1;
      uint64_t j = i + 1;
      if (p.remove_pcr_duplicates)
        for (; j < n && same(recs[j], recs[i]); ++j) { ++dups; if (recs[j].mapq > keep.mapq) keep = recs[j]; }
      if (keep.mapq >= p.mapq_threshold) {
        keep.num_dups = (uint8_t)std::min<uint32_t>(255, dups);
        if (p.tn5_shift) tn5(keep);
        recs[o++] = keep;
      }
      i = j;
    }
    *n_out = o;
    return CMX_OK;
  }
  if (p.tn5_shift) for (uint64_t i = 0; i < n; ++i) tn5(recs[i]);  // chromap.h:1322-1355
  std::sort(recs, recs + n, less);
  if (p.remove_pcr_duplicates) {  // mapping_processor.h:161-202: keeps the last of each run
    uint64_t i = 0, w = 0;
    while (i < n) {
      uint64_t j = i + 1;
      while (j < n && same(recs[j], recs[i])) ++j;
      cmx_pe_record keep = recs[j - 1];
      keep.num_dups = (uint8_t)std::min<uint64_t>(255, j - i);
      recs[w++] = keep;
      i = j;
    }
    n = w;
  }
  for (uint64_t i = 0; i < n; ++i) if (recs[i].mapq >= p.mapq_threshold) recs[o++] = recs[i];
  *n_out = o;
  return CMX_OK;
}

int cmx_postprocess_pairs(cmx_ctx *ctx, cmx_pairs_record *recs, uint64_t n, uint64_t *n_out) {
  if (!ctx || (!recs && n) || !n_out) return CMX_ERR_INVALID;
  const cmx_params &p = ctx->params;
  // records live in the bucket of rid1; the merge order is (bucket, PairsMapping::operator<) (pairs_mapping.h:40-43)
  std::sort(recs, recs + n, [](const cmx_pairs_record &a, const cmx_pairs_record &b) {
    return std::make_tuple(a.rid1, a.rid2, a.pos1, a.pos2, a.mapq, a.read_id) < std::make_tuple(b.rid1, b.rid2, b.pos1, b.pos2, b.mapq, b.read_id);
  });
  uint64_t o = 0;
  if (p.remove_pcr_duplicates) {  // mapping_writer.h:234-300 with PairsMapping::operator== (pairs_mapping.h:44-49)
    uint64_t i = 0;
    while (i < n) {
      cmx_pairs_record keep = recs[i];
      uint64_t j = i + 1;
      for (; j < n && recs[j].rid1 == recs[i].rid1 && recs[j].pos1 == recs[i].pos1 && recs[j].rid2 == recs[i].rid2 && recs[j].pos2 == recs[i].pos2; ++j)
        if (recs[j].mapq > keep.mapq) keep = recs[j];
      if (keep.mapq >= p.mapq_threshold) recs[o++] = keep;
      i = j;
    }
  } else {
    for (uint64_t i = 0; i < n; ++i) if (recs[i].mapq >= p.mapq_threshold) recs[o++] = recs[i];
  }
  *n_out = o;
  return CMX_OK;
}

int64_t cmx_format_pairs(const char *const *names, const uint32_t *lengths, uint32_t n_seq, const cmx_pairs_record *recs, uint64_t n,
                         const char *const *read_names, uint32_t first_read_id, char *buf, int64_t cap) {
  int64_t len = 0;
  std::string hdr = "## pairs format v1.0.0\n#shape: upper triangle\n";  // mapping_writer.cc:383-402
  for (uint32_t i = 0; i < n_seq; ++i) hdr += std::string("#chromsize: ") + names[i] + " " + std::to_string(lengths[i]) + "\n";
  hdr += "#columns: readID chrom1 pos1 chrom2 pos2 strand1 strand2 pair_type mapq1 mapq2\n";
  if (buf && (int64_t)hdr.size() <= cap) memcpy(buf, hdr.data(), hdr.size());
  len += (int64_t)hdr.size();
  for (uint64_t i = 0; i < n; ++i) {  // mapping_writer.cc:405-421
    const cmx_pairs_record &r = recs[i];
    const std::string line = std::string(read_names[r.read_id - first_read_id]) + "\t" + names[r.rid1] + "\t" + std::to_string(r.pos1 + 1) + "\t" + names[r.rid2] +
                             "\t" + std::to_string(r.pos2 + 1) + "\t" + (r.strand1 ? "+" : "-") + "\t" + (r.strand2 ? "+" : "-") + "\tUU\t" +
                             std::to_string(r.mapq) + "\t" + std::to_string(r.mapq) + "\n";
    if (buf && len + (int64_t)line.size() <= cap) memcpy(buf + len, line.data(), line.size());
    len += (int64_t)line.size();
  }
  return len;
}

int cmx_postprocess_bc(cmx_ctx *ctx, cmx_pe_record *recs, uint64_t *bcs, uint64_t n, uint64_t *n_out) {
  if (!ctx || (n && (!recs || !bcs)) || !n_out) return CMX_ERR_INVALID;
  const cmx_params &p = ctx->params;
  *n_out = 0;
  if (n == 0) return CMX_OK;
  std::vector<cmx_pe_record> rr(recs, recs + n);
  std::vector<uint64_t> bb(bcs, bcs + n);
  if (!p.low_memory_mode && p.tn5_shift) for (auto &r : rr) { if (p.single_end) tn5_se(r); else tn5(r); }
  std::vector<uint64_t> ord(n);
  for (uint64_t i = 0; i < n; ++i) ord[i] = i;
  auto key = [&](uint64_t i) {  // bed_mapping.h:145-153 prefixed by rid
    const cmx_pe_record &r = rr[i];
    return std::make_tuple(r.rid, r.fragment_start, r.fragment_length, bb[i], r.mapq, r.direction, r.is_unique, r.read_id);
  };
  std::sort(ord.begin(), ord.end(), [&](uint64_t a, uint64_t b) { return key(a) < key(b); });
  const bool se = p.single_end != 0;
  auto same = [&](uint64_t a, uint64_t b) {  // cell-level duplicates: bed_mapping.h:154-159 (paired-end), :36-39 (single-end: barcode + start)
    return rr[a].rid == rr[b].rid && rr[a].fragment_start == rr[b].fragment_start && (se || rr[a].fragment_length == rr[b].fragment_length) && bb[a] == bb[b];
  };
  uint64_t o = 0, i = 0;
  while (i < n) {
    uint64_t j = i + 1, keep = ord[i];
    uint32_t dups = 1;
    if (p.remove_pcr_duplicates)
      for (; j < n && same(ord[j], ord[j - 1]); ++j) {  // consecutive equality (single-end: the key is not a prefix of the order)
        ++dups;
        if (p.low_memory_mode) { if (rr[ord[j]].mapq > rr[keep].mapq) keep = ord[j]; }
        else keep = ord[j];
      }
    cmx_pe_record k = rr[keep];
    if (k.mapq >= p.mapq_threshold) {
      if (p.remove_pcr_duplicates) k.num_dups = (uint8_t)std::min<uint32_t>(255, dups);
      if (p.low_memory_mode && p.tn5_shift) { if (se) tn5_se(k); else tn5(k); }
      recs[o] = k; bcs[o] = bb[keep]; ++o;
    }
    i = j;
  }
  *n_out = o;
  return CMX_OK;
}

int64_t cmx_format_bed_bc(const char *const *names, const cmx_pe_record *recs, const uint64_t *bcs, uint64_t n, uint32_t bc_len, char *buf, int64_t cap) {
  int64_t len = 0;
  static const char tab[4] = {'A', 'C', 'G', 'T'};
  for (uint64_t i = 0; i < n; ++i) {  // mapping_writer.cc:127-137, barcode_translator.h:114-123
    const cmx_pe_record &r = recs[i];
    std::string line = std::string(names[r.rid]) + "\t" + std::to_string(r.fragment_start) + "\t" + std::to_string((uint32_t)(r.fragment_start + r.fragment_length)) + "\t";
    for (uint32_t j = 0; j < bc_len; ++j) line.push_back(tab[(bcs[i] >> ((bc_len - 1 - j) * 2)) & 3]);
    line += "\t" + std::to_string((uint32_t)r.num_dups) + "\n";
    if (buf && len + (int64_t)line.size() <= cap) memcpy(buf + len, line.data(), line.size());
    len += (int64_t)line.size();
  }
  return len;
}

// Sort + duplicate removal + MAPQ filter (+ Tn5 shift) over device-resident records: d_a[0, n) (and d_bca) in, the result in
// d_b[0, *nsel) (and d_bcb); d_a is used as scratch.  All four buffers hold n entries.  Work is queued on ctx->stream and
// waited for.
static int pp_device(cmx_ctx *ctx, const PpParams &P, PpRecord *d_a, u64 *d_bca, u64 n, PpRecord *d_b, u64 *d_bcb, u64 *nsel_out) {
  *nsel_out = 0;
  if (n == 0) return CMX_OK;
  if (n > 0x7FFFFFFFull) return fail(ctx, CMX_ERR_INVALID, "post-processing: more than 2^31-1 records in one call");
  const bool bc = P.kind == PP_BED_BC;
  cudaStream_t st = ctx->stream;
  u64 *d_k0 = nullptr, *d_k1 = nullptr, *d_nsel = nullptr;
  u32 *d_i0 = nullptr, *d_i1 = nullptr;
  u8 *d_head = nullptr, *d_keep = nullptr;
  void *d_tmp = nullptr;
  auto cleanup = [&]() { cudaFree(d_k0); cudaFree(d_k1); cudaFree(d_nsel); cudaFree(d_i0); cudaFree(d_i1); cudaFree(d_head); cudaFree(d_keep); cudaFree(d_tmp); };
#define PPCU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); return fail(ctx, CMX_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); } } while (0)
  PPCU(cudaMalloc(&d_k0, n * 8)); PPCU(cudaMalloc(&d_k1, n * 8)); PPCU(cudaMalloc(&d_i0, n * 4)); PPCU(cudaMalloc(&d_i1, n * 4));
  PPCU(cudaMalloc(&d_head, n)); PPCU(cudaMalloc(&d_keep, n)); PPCU(cudaMalloc(&d_nsel, 16));
  const unsigned nb = (unsigned)((n + 255) / 256);
  if (!P.low_mem && P.tn5 && P.kind != PP_PAIRS) pp_tn5_kernel<<<nb, 256, 0, st>>>(P.kind, P.se, d_a, n);  // chromap.h:1322-1355: before the sort
  pp_iota_kernel<<<nb, 256, 0, st>>>(d_i0, n);
  cub::DoubleBuffer<u64> dk(d_k0, d_k1);
  cub::DoubleBuffer<u32> di(d_i0, d_i1);
  size_t tmp_bytes = 0, need = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, di, (int)n, 0, 64, st);
  cub::DeviceSelect::Flagged(nullptr, need, d_a, d_keep, d_b, d_nsel, (int)n, st);
  tmp_bytes = std::max(tmp_bytes, need);
  cub::DeviceSelect::Flagged(nullptr, need, d_bca, d_keep, d_bcb, d_nsel, (int)n, st);
  tmp_bytes = std::max(tmp_bytes, need);
  PPCU(cudaMalloc(&d_tmp, tmp_bytes));
  // only the bits a key word can hold are sorted: word 0 (rid | start, or rid1 | rid2) up to its largest value in this call;
  // pp_key_word bounds the others: the 32-bit alignment lengths; length alone; mapq | direction | unique | read id
  u64 max0 = 0;
  PPCU(cudaMemsetAsync(d_nsel, 0, 16, st));
  pp_max_key0_kernel<<<std::min(nb, 1184u), 256, 0, st>>>(P.kind, d_a, n, d_nsel);
  PPCU(cudaMemcpyAsync(&max0, d_nsel, 8, cudaMemcpyDeviceToHost, st));
  PPCU(cudaStreamSynchronize(st));
  int bits0 = 1;
  while (bits0 < 64 && (max0 >> bits0)) ++bits0;
  for (int w = pp_n_words(P.kind) - 1; w >= 0; --w) {  // least significant word first; every pass is stable
    pp_key_kernel<<<nb, 256, 0, st>>>(P.kind, w, d_a, bc ? d_bca : nullptr, di.Current(), n, dk.Current());
    int bits = 64;
    if (w == 0) bits = bits0;
    else if (P.kind == PP_PAIRS) bits = w == 2 ? 40 : 64;
    else if (bc) bits = w == 1 ? 16 : (w == 3 ? 56 : 64);
    else if (w == 2) bits = 32;
    PPCU(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, dk, di, (int)n, 0, bits, st));
  }
  pp_gather_kernel<<<nb, 256, 0, st>>>(d_a, bc ? d_bca : nullptr, di.Current(), n, d_b, d_bcb);
  pp_head_kernel<<<nb, 256, 0, st>>>(P.kind, P.se, P.dedup, d_b, bc ? d_bcb : nullptr, n, d_head);
  pp_resolve_kernel<<<nb, 256, 0, st>>>(P, d_b, bc ? d_bcb : nullptr, d_head, n, d_a, bc ? d_bca : nullptr, d_keep);
  PPCU(cub::DeviceSelect::Flagged(d_tmp, tmp_bytes, d_a, d_keep, d_b, d_nsel, (int)n, st));
  if (bc) PPCU(cub::DeviceSelect::Flagged(d_tmp, tmp_bytes, d_bca, d_keep, d_bcb, d_nsel + 1, (int)n, st));
  u64 nsel = 0;
  PPCU(cudaMemcpyAsync(&nsel, d_nsel, 8, cudaMemcpyDeviceToHost, st));
  PPCU(cudaStreamSynchronize(st));
  PPCU(cudaGetLastError());
#undef PPCU
  cleanup();
  *nsel_out = nsel;
  return CMX_OK;
}

// The same in place on host buffers: same results as cmx_postprocess / cmx_postprocess_bc / cmx_postprocess_pairs
// (postprocess.cuh).  The record kind follows the context: pairs when output_format == 5, barcoded BED when
// barcode_keys != NULL, else bulk BED.
int cmx_postprocess_gpu(cmx_ctx *ctx, void *records, uint64_t *barcode_keys, uint64_t n, uint64_t *n_out) {
  if (!ctx || (!records && n) || !n_out) return CMX_ERR_INVALID;
  *n_out = 0;
  if (n == 0) return CMX_OK;
  if (n > 0x7FFFFFFFull) return fail(ctx, CMX_ERR_INVALID, "cmx_postprocess_gpu: more than 2^31-1 records in one call");
  CU(cudaSetDevice(ctx->device));
  const cmx_params &p = ctx->params;
  PpParams P;
  P.kind = p.output_format == 5 ? PP_PAIRS : (barcode_keys ? PP_BED_BC : (p.single_end ? PP_BED_SE : PP_BED));
  P.low_mem = p.low_memory_mode; P.dedup = p.remove_pcr_duplicates; P.tn5 = p.tn5_shift; P.mapq_threshold = p.mapq_threshold; P.se = p.single_end;
  if (P.kind == PP_PAIRS && barcode_keys) return fail(ctx, CMX_ERR_INVALID, "barcodes are not supported with pairs output");
  const bool bc = P.kind == PP_BED_BC;
  cudaStream_t st = ctx->stream;
  PpRecord *d_a = nullptr, *d_b = nullptr;
  u64 *d_bca = nullptr, *d_bcb = nullptr;
  auto cleanup = [&]() { cudaFree(d_a); cudaFree(d_b); cudaFree(d_bca); cudaFree(d_bcb); };
#define PPCU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); return fail(ctx, CMX_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); } } while (0)
  PPCU(cudaMalloc(&d_a, n * sizeof(PpRecord))); PPCU(cudaMalloc(&d_b, n * sizeof(PpRecord)));
  if (bc) { PPCU(cudaMalloc(&d_bca, n * 8)); PPCU(cudaMalloc(&d_bcb, n * 8)); }
  PPCU(cudaMemcpyAsync(d_a, records, n * sizeof(PpRecord), cudaMemcpyHostToDevice, st));
  if (bc) PPCU(cudaMemcpyAsync(d_bca, barcode_keys, n * 8, cudaMemcpyHostToDevice, st));
  u64 nsel = 0;
  const int rc = pp_device(ctx, P, d_a, d_bca, n, d_b, d_bcb, &nsel);
  if (rc != CMX_OK) { cleanup(); return rc; }
  if (nsel) PPCU(cudaMemcpyAsync(records, d_b, nsel * sizeof(PpRecord), cudaMemcpyDeviceToHost, st));
  if (bc && nsel) PPCU(cudaMemcpyAsync(barcode_keys, d_bcb, nsel * 8, cudaMemcpyDeviceToHost, st));
  PPCU(cudaStreamSynchronize(st));
#undef PPCU
  cleanup();
  *n_out = nsel;
  return CMX_OK;
}

// ---- multi-GPU duplicate-removal exchange (exchange.cuh) -------------------------------------------------
// NCCL is reached through dlopen: the library has no link-time dependency on it, and inside a process that already
// holds an NCCL (PyTorch's) that copy is the one used.  Types are declared here (ABI of nccl.h 2.x).
namespace {
struct NcclApi {
  typedef struct { char internal[128]; } UniqueId;
  int (*GetUniqueId)(UniqueId *) = nullptr;
  int (*CommInitRank)(void **, int, UniqueId, int) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
  int (*Send)(const void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  bool ok = false;
  std::string why;
};
NcclApi &nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  // an NCCL the process already holds (PyTorch's) first; otherwise CMX_NCCL_LIB, then the system library.  RTLD_LOCAL: our copy
  // must not satisfy the symbol lookups of a framework that is loaded later and expects its own (newer) NCCL.
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_LOCAL);
  if (!h) { const char *pth = getenv("CMX_NCCL_LIB"); if (pth && *pth) h = dlopen(pth, RTLD_NOW | RTLD_LOCAL); }
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
  if (!h) { api.why = std::string("libnccl.so.2 not found: ") + dlerror(); return api; }
  api.GetUniqueId = (int (*)(NcclApi::UniqueId *))dlsym(h, "ncclGetUniqueId");
  api.CommInitRank = (int (*)(void **, int, NcclApi::UniqueId, int))dlsym(h, "ncclCommInitRank");
  api.CommDestroy = (int (*)(void *))dlsym(h, "ncclCommDestroy");
  api.AllGather = (int (*)(const void *, void *, size_t, int, void *, cudaStream_t))dlsym(h, "ncclAllGather");
  api.Send = (int (*)(const void *, size_t, int, int, void *, cudaStream_t))dlsym(h, "ncclSend");
  api.Recv = (int (*)(void *, size_t, int, int, void *, cudaStream_t))dlsym(h, "ncclRecv");
  api.GroupStart = (int (*)())dlsym(h, "ncclGroupStart");
  api.GroupEnd = (int (*)())dlsym(h, "ncclGroupEnd");
  api.GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
  api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.Send && api.Recv && api.GroupStart && api.GroupEnd;
  if (!api.ok) api.why = "libnccl.so.2 lacks the expected symbols";
  return api;
}
const int NCCL_UINT8 = 1, NCCL_UINT64 = 5;  // ncclDataType_t
}  // namespace

int cmx_comm_unique_id(void *id128) {
  if (!id128) return CMX_ERR_INVALID;
  NcclApi &N = nccl_api();
  if (!N.ok) return CMX_ERR_STATE;
  NcclApi::UniqueId id;
  if (N.GetUniqueId(&id) != 0) return CMX_ERR_CUDA;
  memcpy(id128, &id, 128);
  return CMX_OK;
}

int cmx_comm_init(cmx_ctx *ctx, int n_ranks, int rank, const void *id128) {
  if (!ctx || !id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) return CMX_ERR_INVALID;
  NcclApi &N = nccl_api();
  if (!N.ok) return fail(ctx, CMX_ERR_STATE, "NCCL unavailable: %s", N.why.c_str());
  CU(cudaSetDevice(ctx->device));
  if (ctx->nccl_comm) { N.CommDestroy(ctx->nccl_comm); ctx->nccl_comm = nullptr; }
  NcclApi::UniqueId id;
  memcpy(&id, id128, 128);
  const int rc = N.CommInitRank(&ctx->nccl_comm, n_ranks, id, rank);
  if (rc != 0) return fail(ctx, CMX_ERR_CUDA, "ncclCommInitRank: %s", N.GetErrorString ? N.GetErrorString(rc) : "error");
  ctx->comm_rank = rank; ctx->comm_size = n_ranks;
  return CMX_OK;
}

int cmx_comm_destroy(cmx_ctx *ctx) {
  if (!ctx) return CMX_ERR_INVALID;
  if (ctx->nccl_comm) { nccl_api().CommDestroy(ctx->nccl_comm); ctx->nccl_comm = nullptr; }
  ctx->comm_rank = 0; ctx->comm_size = 1;
  return CMX_OK;
}

// This rank's records in, this rank's survivors out (reference order, duplicate counts set, MAPQ-filtered, Tn5 NOT yet
// applied: the low-memory merge shifts after the final ordering, mapping_writer.h:285-287 — cmx_exchange_finish below).
int cmx_dedup_exchange(cmx_ctx *ctx, const void *records, const uint64_t *barcode_keys, uint64_t n, int on_device, void *out_records,
                       uint64_t *out_barcode_keys, uint64_t *n_out, cmx_exchange_stats *stats) {
  if (!ctx || (!records && n) || !out_records || !n_out) return CMX_ERR_INVALID;
  *n_out = 0;
  if (stats) memset(stats, 0, sizeof(*stats));
  const cmx_params &p = ctx->params;
  if (p.output_format == 5 || p.single_end || !p.low_memory_mode)
    return fail(ctx, CMX_ERR_INVALID, "cmx_dedup_exchange: paired-end BED records in low-memory mode only (every preset with duplicate removal)");
  if (!ctx->nccl_comm) return fail(ctx, CMX_ERR_STATE, "cmx_dedup_exchange: cmx_comm_init first");
  if (n >= 0x7FFFFFFFull) return fail(ctx, CMX_ERR_INVALID, "cmx_dedup_exchange: more than 2^31-1 records on one rank");
  NcclApi &N = nccl_api();
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const bool bc = barcode_keys != nullptr;
  const int tw = bc ? 3 : 2, R = ctx->comm_size, rank = ctx->comm_rank;
  PpRecord *d_rec = nullptr, *d_out = nullptr;
  u64 *d_bc = nullptr, *d_outbc = nullptr, *d_cnt = nullptr, *d_send = nullptr, *d_all = nullptr, *d_k0 = nullptr, *d_k1 = nullptr, *d_nsel = nullptr;
  u32 *d_i0 = nullptr, *d_i1 = nullptr, *d_sel = nullptr, *d_selc = nullptr;
  u8 *d_head = nullptr, *d_keep = nullptr, *d_dups = nullptr, *d_dupsc = nullptr;
  void *d_tmp = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  auto cleanup = [&]() {
    if (!on_device) { cudaFree(d_rec); cudaFree(d_bc); cudaFree(d_out); cudaFree(d_outbc); }
    cudaFree(d_cnt); cudaFree(d_send); cudaFree(d_all); cudaFree(d_k0); cudaFree(d_k1); cudaFree(d_nsel); cudaFree(d_i0); cudaFree(d_i1);
    cudaFree(d_sel); cudaFree(d_selc); cudaFree(d_head); cudaFree(d_keep); cudaFree(d_dups); cudaFree(d_dupsc); cudaFree(d_tmp);
    for (auto &e : ev) if (e) cudaEventDestroy(e);
  };
#define XCU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); return fail(ctx, CMX_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); } } while (0)
#define XNC(call) do { int r_ = (call); if (r_ != 0) { cleanup(); return fail(ctx, CMX_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, N.GetErrorString ? N.GetErrorString(r_) : "NCCL error"); } } while (0)
  for (auto &e : ev) XCU(cudaEventCreate(&e));
  if (on_device) { d_rec = (PpRecord *)records; d_bc = (u64 *)barcode_keys; d_out = (PpRecord *)out_records; d_outbc = (u64 *)out_barcode_keys; }
  else {
    XCU(cudaMalloc(&d_rec, std::max<u64>(n, 1) * sizeof(PpRecord))); XCU(cudaMalloc(&d_out, std::max<u64>(n, 1) * sizeof(PpRecord)));
    XCU(cudaMemcpyAsync(d_rec, records, n * sizeof(PpRecord), cudaMemcpyHostToDevice, st));
    if (bc) { XCU(cudaMalloc(&d_bc, std::max<u64>(n, 1) * 8)); XCU(cudaMalloc(&d_outbc, std::max<u64>(n, 1) * 8)); XCU(cudaMemcpyAsync(d_bc, barcode_keys, n * 8, cudaMemcpyHostToDevice, st)); }
  }
  // sizes: an 8-byte all-gather of the record counts
  XCU(cudaMalloc(&d_cnt, (size_t)(R + 1) * 8));
  XCU(cudaMemcpyAsync(d_cnt + R, &n, 8, cudaMemcpyHostToDevice, st));
  XNC(N.AllGather(d_cnt + R, d_cnt, 1, NCCL_UINT64, ctx->nccl_comm, st));
  std::vector<u64> cnt(R);
  XCU(cudaMemcpyAsync(cnt.data(), d_cnt, (size_t)R * 8, cudaMemcpyDeviceToHost, st));
  XCU(cudaStreamSynchronize(st));
  u64 n_pad = 1, n_total = 0;
  for (u64 c : cnt) { n_pad = std::max(n_pad, c); n_total += c; }
  const u64 n_all = n_pad * (u64)R;
  if (n_all >= 0xFFFFFFFFull) { cleanup(); return fail(ctx, CMX_ERR_INVALID, "cmx_dedup_exchange: %llu gathered tuples exceed 2^32", (unsigned long long)n_all); }
  // pack -> ONE all-gather of the tuples -> sort -> decide
  XCU(cudaMalloc(&d_send, n_pad * tw * 8)); XCU(cudaMalloc(&d_all, n_all * tw * 8));
  XCU(cudaMalloc(&d_k0, n_all * 8)); XCU(cudaMalloc(&d_k1, n_all * 8)); XCU(cudaMalloc(&d_i0, n_all * 4)); XCU(cudaMalloc(&d_i1, n_all * 4));
  XCU(cudaMalloc(&d_head, n_all)); XCU(cudaMalloc(&d_keep, n_all)); XCU(cudaMalloc(&d_dups, n_all)); XCU(cudaMalloc(&d_dupsc, std::max<u64>(n, 1)));
  XCU(cudaMalloc(&d_sel, n_all * 4)); XCU(cudaMalloc(&d_selc, std::max<u64>(n, 1) * 4)); XCU(cudaMalloc(&d_nsel, 16));
  XCU(cudaEventRecord(ev[0], st));
  ex_pack_kernel<<<(unsigned)((n_pad + 255) / 256), 256, 0, st>>>(d_rec, d_bc, n, n_pad, bc ? 1 : 0, d_send);
  XCU(cudaEventRecord(ev[1], st));
  XNC(N.AllGather(d_send, d_all, n_pad * tw * 8, NCCL_UINT8, ctx->nccl_comm, st));
  XCU(cudaEventRecord(ev[2], st));
  const unsigned nb = (unsigned)((n_all + 255) / 256);
  pp_iota_kernel<<<nb, 256, 0, st>>>(d_i0, n_all);
  cub::DoubleBuffer<u64> dk(d_k0, d_k1);
  cub::DoubleBuffer<u32> di(d_i0, d_i1);
  size_t tmp_bytes = 0, need = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, di, (int)n_all, 0, 64, st);
  cub::DeviceSelect::Flagged(nullptr, need, d_sel, d_keep, d_selc, d_nsel, (int)n_all, st);
  tmp_bytes = std::max(tmp_bytes, need);
  XCU(cudaMalloc(&d_tmp, tmp_bytes));
  // least significant key first, every pass stable: bulk (b, a); barcoded (low 48 bits of b, barcode, length, a)
  const int passes_bulk[2] = {4, 3}, passes_bc[4] = {0, 1, 2, 3};
  for (int q = 0; q < (bc ? 4 : 2); ++q) {
    const int pass = bc ? passes_bc[q] : passes_bulk[q];
    ex_key_kernel<<<nb, 256, 0, st>>>(d_all, tw, pass, di.Current(), n_all, dk.Current());
    XCU(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, dk, di, (int)n_all, 0, pass == 2 ? 16 : 64, st));
  }
  // padding sorts last (a = ~0): only the first n_total sorted entries are records
  if (n_total) {
    const unsigned nbt = (unsigned)((n_total + 255) / 256);
    ex_head_kernel<<<nbt, 256, 0, st>>>(d_all, tw, p.remove_pcr_duplicates, di.Current(), n_total, d_head);
    ex_resolve_kernel<<<nbt, 256, 0, st>>>(d_all, tw, di.Current(), d_head, n_total, n_pad, rank, p.mapq_threshold, d_keep, d_sel, d_dups);
    XCU(cub::DeviceSelect::Flagged(d_tmp, tmp_bytes, d_sel, d_keep, d_selc, d_nsel, (int)n_total, st));
    XCU(cub::DeviceSelect::Flagged(d_tmp, tmp_bytes, d_dups, d_keep, d_dupsc, d_nsel + 1, (int)n_total, st));
  } else XCU(cudaMemsetAsync(d_nsel, 0, 16, st));
  u64 nsel = 0;
  XCU(cudaMemcpyAsync(&nsel, d_nsel, 8, cudaMemcpyDeviceToHost, st));
  XCU(cudaStreamSynchronize(st));
  if (nsel) ex_gather_kernel<<<(unsigned)((nsel + 255) / 256), 256, 0, st>>>(d_rec, d_bc, d_selc, d_dupsc, p.remove_pcr_duplicates, nsel, d_out, bc ? d_outbc : nullptr);
  XCU(cudaEventRecord(ev[3], st));
  if (!on_device && nsel) {
    XCU(cudaMemcpyAsync(out_records, d_out, nsel * sizeof(PpRecord), cudaMemcpyDeviceToHost, st));
    if (bc && out_barcode_keys) XCU(cudaMemcpyAsync(out_barcode_keys, d_outbc, nsel * 8, cudaMemcpyDeviceToHost, st));
  }
  XCU(cudaStreamSynchronize(st));
  XCU(cudaGetLastError());
  if (stats) {
    cudaEventElapsedTime(&stats->pack_ms, ev[0], ev[1]);
    cudaEventElapsedTime(&stats->allgather_ms, ev[1], ev[2]);
    cudaEventElapsedTime(&stats->resolve_ms, ev[2], ev[3]);
    stats->bytes_sent = n_pad * tw * 8; stats->bytes_received = n_all * tw * 8; stats->n_global = n_total; stats->n_ranks = (uint32_t)R;
  }
#undef XCU
#undef XNC
  cleanup();
  *n_out = nsel;
  return CMX_OK;
}

// The same step as a range shuffle (exchange.cuh, second half): this rank's records in; out = the records of THIS RANK'S KEY
// RANGE after duplicate removal over the whole run, in the reference's order, num_dups set, MAPQ-filtered, Tn5 applied —
// the run's output is the ranks' outputs one after the other in rank order.  Work per rank is proportional to its share of
// the run (the all-gather variant above sorts every rank's tuples on every rank).
int cmx_dedup_shuffle(cmx_ctx *ctx, const void *records, const uint64_t *barcode_keys, uint64_t n, int on_device, void *out_records,
                      uint64_t *out_barcode_keys, uint64_t out_capacity, uint64_t *n_out, cmx_shuffle_stats *stats) {
  if (!ctx || (!records && n) || (!out_records && out_capacity) || !n_out) return CMX_ERR_INVALID;
  *n_out = 0;
  if (stats) memset(stats, 0, sizeof(*stats));
  const cmx_params &p = ctx->params;
  if (p.output_format == 5 || p.single_end || !p.low_memory_mode)
    return fail(ctx, CMX_ERR_INVALID, "cmx_dedup_shuffle: paired-end BED records in low-memory mode only (every preset with duplicate removal)");
  if (!ctx->nccl_comm) return fail(ctx, CMX_ERR_STATE, "cmx_dedup_shuffle: cmx_comm_init first");
  if (n >= 0x7FFFFFFFull) return fail(ctx, CMX_ERR_INVALID, "cmx_dedup_shuffle: more than 2^31-1 records on one rank");
  NcclApi &N = nccl_api();
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const bool bc = barcode_keys != nullptr;
  const int R = ctx->comm_size, rank = ctx->comm_rank;
  PpRecord *d_rec = nullptr, *d_send = nullptr, *d_recv = nullptr, *d_res = nullptr;
  u64 *d_bc = nullptr, *d_sendbc = nullptr, *d_recvbc = nullptr, *d_resbc = nullptr, *d_a = nullptr, *d_samp = nullptr, *d_all0 = nullptr, *d_all1 = nullptr,
      *d_split = nullptr, *d_off = nullptr, *d_offall = nullptr;
  u32 *d_d0 = nullptr, *d_d1 = nullptr, *d_i0 = nullptr, *d_i1 = nullptr;
  void *d_tmp = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  auto cleanup = [&]() {
    if (!on_device) { cudaFree(d_rec); cudaFree(d_bc); }
    cudaFree(d_send); cudaFree(d_recv); cudaFree(d_res); cudaFree(d_sendbc); cudaFree(d_recvbc); cudaFree(d_resbc); cudaFree(d_a); cudaFree(d_samp);
    cudaFree(d_all0); cudaFree(d_all1); cudaFree(d_split); cudaFree(d_off); cudaFree(d_offall); cudaFree(d_d0); cudaFree(d_d1); cudaFree(d_i0); cudaFree(d_i1);
    cudaFree(d_tmp);
    for (auto &e : ev) if (e) cudaEventDestroy(e);
  };
#define XCU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); return fail(ctx, CMX_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); } } while (0)
#define XNC(call) do { int r_ = (call); if (r_ != 0) { cleanup(); return fail(ctx, CMX_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, N.GetErrorString ? N.GetErrorString(r_) : "NCCL error"); } } while (0)
  for (auto &e : ev) XCU(cudaEventCreate(&e));
  const u64 n1 = std::max<u64>(n, 1);
  if (on_device) { d_rec = (PpRecord *)records; d_bc = (u64 *)barcode_keys; }
  else {
    XCU(cudaMalloc(&d_rec, n1 * sizeof(PpRecord)));
    XCU(cudaMemcpyAsync(d_rec, records, n * sizeof(PpRecord), cudaMemcpyHostToDevice, st));
    if (bc) { XCU(cudaMalloc(&d_bc, n1 * 8)); XCU(cudaMemcpyAsync(d_bc, barcode_keys, n * 8, cudaMemcpyHostToDevice, st)); }
  }
  const u64 n_samp = (u64)R * SH_SAMPLE;
  XCU(cudaMalloc(&d_a, n1 * 8)); XCU(cudaMalloc(&d_samp, SH_SAMPLE * 8)); XCU(cudaMalloc(&d_all0, n_samp * 8)); XCU(cudaMalloc(&d_all1, n_samp * 8));
  XCU(cudaMalloc(&d_split, (size_t)std::max(R - 1, 1) * 8)); XCU(cudaMalloc(&d_off, (size_t)(R + 1) * 8)); XCU(cudaMalloc(&d_offall, (size_t)R * (R + 1) * 8));
  XCU(cudaMalloc(&d_d0, n1 * 4)); XCU(cudaMalloc(&d_d1, n1 * 4)); XCU(cudaMalloc(&d_i0, n1 * 4)); XCU(cudaMalloc(&d_i1, n1 * 4));
  XCU(cudaMalloc(&d_send, n1 * sizeof(PpRecord)));
  if (bc) XCU(cudaMalloc(&d_sendbc, n1 * 8));
  size_t tmp_bytes = 0, need = 0;
  {
    cub::DoubleBuffer<u64> ks(d_all0, d_all1);
    cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, ks, (int)n_samp, 0, 64, st);
    cub::DoubleBuffer<u32> dd(d_d0, d_d1), di(d_i0, d_i1);
    cub::DeviceRadixSort::SortPairs(nullptr, need, dd, di, (int)n1, 0, 32, st);
    tmp_bytes = std::max(tmp_bytes, need);
  }
  XCU(cudaMalloc(&d_tmp, tmp_bytes));
  // ---- partition: splitters from an all-gathered sample, destination of every record, records grouped by destination
  XCU(cudaEventRecord(ev[0], st));
  const unsigned nb = (unsigned)((n1 + 255) / 256);
  if (n) sh_key_kernel<<<nb, 256, 0, st>>>(d_rec, n, d_a);
  sh_sample_kernel<<<(SH_SAMPLE + 255) / 256, 256, 0, st>>>(d_a, n, d_samp);
  XNC(N.AllGather(d_samp, d_all0, SH_SAMPLE, NCCL_UINT64, ctx->nccl_comm, st));
  cub::DoubleBuffer<u64> ks(d_all0, d_all1);
  XCU(cub::DeviceRadixSort::SortKeys(d_tmp, tmp_bytes, ks, (int)n_samp, 0, 64, st));
  std::vector<u64> samp(n_samp), split(std::max(R - 1, 1), 0);
  XCU(cudaMemcpyAsync(samp.data(), ks.Current(), n_samp * 8, cudaMemcpyDeviceToHost, st));
  XCU(cudaStreamSynchronize(st));
  const u64 valid = (u64)(std::lower_bound(samp.begin(), samp.end(), (u64)EX_PAD) - samp.begin());  // padding sorts last
  for (int j = 1; j < R; ++j) split[j - 1] = valid ? samp[valid * (u64)j / (u64)R] : 0ull;           // the same on every rank
  XCU(cudaMemcpyAsync(d_split, split.data(), split.size() * 8, cudaMemcpyHostToDevice, st));
  int dest_bits = 1;
  while ((1 << dest_bits) < R) ++dest_bits;
  cub::DoubleBuffer<u32> dd(d_d0, d_d1), di(d_i0, d_i1);
  if (n) {
    sh_dest_kernel<<<nb, 256, 0, st>>>(d_a, n, d_split, R - 1, dd.Current());
    pp_iota_kernel<<<nb, 256, 0, st>>>(di.Current(), n);
    XCU(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, dd, di, (int)n, 0, dest_bits, st));  // stable: mapping order kept inside a destination
    pp_gather_kernel<<<nb, 256, 0, st>>>(d_rec, bc ? d_bc : nullptr, di.Current(), n, d_send, d_sendbc);
  }
  sh_bounds_kernel<<<(R + 1 + 63) / 64, 64, 0, st>>>(dd.Current(), n, R, d_off);
  XNC(N.AllGather(d_off, d_offall, (size_t)(R + 1), NCCL_UINT64, ctx->nccl_comm, st));
  std::vector<u64> offall((size_t)R * (R + 1));
  XCU(cudaMemcpyAsync(offall.data(), d_offall, offall.size() * 8, cudaMemcpyDeviceToHost, st));
  XCU(cudaStreamSynchronize(st));
  auto cnt = [&](int from, int to) { return offall[(size_t)from * (R + 1) + to + 1] - offall[(size_t)from * (R + 1) + to]; };
  u64 n_recv = 0, n_global = 0;
  std::vector<u64> roff(R + 1, 0);
  for (int q = 0; q < R; ++q) { roff[q] = n_recv; n_recv += cnt(q, rank); n_global += offall[(size_t)q * (R + 1) + R]; }
  roff[R] = n_recv;
  if (n_recv >= 0x7FFFFFFFull) { cleanup(); return fail(ctx, CMX_ERR_INVALID, "cmx_dedup_shuffle: %llu records in this rank's key range exceed 2^31-1", (unsigned long long)n_recv); }
  const u64 nr1 = std::max<u64>(n_recv, 1);
  XCU(cudaMalloc(&d_recv, nr1 * sizeof(PpRecord))); XCU(cudaMalloc(&d_res, nr1 * sizeof(PpRecord)));
  if (bc) { XCU(cudaMalloc(&d_recvbc, nr1 * 8)); XCU(cudaMalloc(&d_resbc, nr1 * 8)); }
  // ---- shuffle: every record travels once, to the rank that owns its key range
  XCU(cudaEventRecord(ev[1], st));
  XNC(N.GroupStart());
  for (int q = 0; q < R; ++q) {
    const u64 so = offall[(size_t)rank * (R + 1) + q], sc = cnt(rank, q), rc = cnt(q, rank);
    if (q == rank) {  // this rank's own share stays on the device
      if (sc) {
        XCU(cudaMemcpyAsync(d_recv + roff[q], d_send + so, sc * sizeof(PpRecord), cudaMemcpyDeviceToDevice, st));
        if (bc) XCU(cudaMemcpyAsync(d_recvbc + roff[q], d_sendbc + so, sc * 8, cudaMemcpyDeviceToDevice, st));
      }
      continue;
    }
    if (sc) {
      XNC(N.Send(d_send + so, sc * sizeof(PpRecord), NCCL_UINT8, q, ctx->nccl_comm, st));
      if (bc) XNC(N.Send(d_sendbc + so, sc * 8, NCCL_UINT8, q, ctx->nccl_comm, st));
    }
    if (rc) {
      XNC(N.Recv(d_recv + roff[q], rc * sizeof(PpRecord), NCCL_UINT8, q, ctx->nccl_comm, st));
      if (bc) XNC(N.Recv(d_recvbc + roff[q], rc * 8, NCCL_UINT8, q, ctx->nccl_comm, st));
    }
  }
  XNC(N.GroupEnd());
  XCU(cudaEventRecord(ev[2], st));
  // ---- the ordinary single-GPU post-processing of what arrived
  PpParams P;
  P.kind = bc ? PP_BED_BC : PP_BED; P.low_mem = 1; P.dedup = p.remove_pcr_duplicates; P.tn5 = p.tn5_shift; P.mapq_threshold = p.mapq_threshold; P.se = 0;
  u64 nsel = 0;
  const int prc = pp_device(ctx, P, d_recv, d_recvbc, n_recv, d_res, d_resbc, &nsel);
  if (prc != CMX_OK) { cleanup(); return prc; }
  XCU(cudaEventRecord(ev[3], st));
  *n_out = nsel;
  if (nsel > out_capacity) { cleanup(); return fail(ctx, CMX_ERR_INVALID, "cmx_dedup_shuffle: %llu records for a capacity of %llu", (unsigned long long)nsel, (unsigned long long)out_capacity); }
  if (nsel) {
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    XCU(cudaMemcpyAsync(out_records, d_res, nsel * sizeof(PpRecord), kind, st));
    if (bc && out_barcode_keys) XCU(cudaMemcpyAsync(out_barcode_keys, d_resbc, nsel * 8, kind, st));
  }
  XCU(cudaStreamSynchronize(st));
  XCU(cudaGetLastError());
  if (stats) {
    cudaEventElapsedTime(&stats->partition_ms, ev[0], ev[1]);
    cudaEventElapsedTime(&stats->shuffle_ms, ev[1], ev[2]);
    cudaEventElapsedTime(&stats->postprocess_ms, ev[2], ev[3]);
    const u64 rb = sizeof(PpRecord) + (bc ? 8 : 0);
    stats->bytes_sent = (n - cnt(rank, rank)) * rb; stats->bytes_received = (n_recv - cnt(rank, rank)) * rb;
    stats->n_received = n_recv; stats->n_global = n_global; stats->n_ranks = (uint32_t)R;
  }
#undef XCU
#undef XNC
  cleanup();
  return CMX_OK;
}

// Last step after the survivors of all ranks have been brought together (e.g. on rank 0): reference order and the deferred
// Tn5 shift (mapping_writer.h:285-287).  Host only (no device needed), in place.
int cmx_exchange_finish(const cmx_params *params, cmx_pe_record *recs, uint64_t *bcs, uint64_t n) {
  if (!params || (!recs && n)) return CMX_ERR_INVALID;
  std::vector<u64> order(n);
  for (u64 i = 0; i < n; ++i) order[i] = i;
  auto key = [&](u64 i) { const cmx_pe_record &r = recs[i]; return std::make_tuple(r.rid, r.fragment_start, r.fragment_length, bcs ? bcs[i] : 0ull, r.mapq, r.direction, r.is_unique, r.read_id); };
  std::stable_sort(order.begin(), order.end(), [&](u64 x, u64 y) { return key(x) < key(y); });
  std::vector<cmx_pe_record> tmp(n);
  std::vector<u64> tb(bcs ? n : 0);
  for (u64 i = 0; i < n; ++i) { tmp[i] = recs[order[i]]; if (bcs) tb[i] = bcs[order[i]]; }
  for (u64 i = 0; i < n; ++i) { recs[i] = tmp[i]; if (bcs) bcs[i] = tb[i]; if (params->tn5_shift) tn5(recs[i]); }
  return CMX_OK;
}

// BED text on the device (bed_len_kernel / bed_write_kernel), byte-identical to cmx_format_bed / cmx_format_bed_bc.
// Records (and barcode keys, NULL for bulk data) and the output buffer are host memory; work goes through the device in
// chunks of 8 M records.  buf == NULL: returns the text length only.  Returns < 0 on error.
int64_t cmx_format_bed_gpu(cmx_ctx *ctx, const char *const *names, const cmx_pe_record *records, const uint64_t *barcode_keys, uint64_t n,
                           uint32_t bc_len, char *buf, int64_t cap) {
  if (!ctx || !names || (!records && n) || (barcode_keys && (bc_len == 0 || bc_len > 32))) return -1;
  if (cudaSetDevice(ctx->device) != cudaSuccess) return -1;
  const u32 n_seq = ctx->n_seq;
  std::string cat;
  std::vector<u32> noff(n_seq + 1, 0);
  for (u32 i = 0; i < n_seq; ++i) { cat += names[i]; noff[i + 1] = (u32)cat.size(); }
  const u64 CH = 8u << 20;
  const u64 nc = std::min<u64>(n, CH);
  char *d_names = nullptr, *d_out = nullptr;
  u32 *d_noff = nullptr, *d_len = nullptr;
  PpRecord *d_rec = nullptr;
  u64 *d_bc = nullptr, *d_off = nullptr;
  void *d_tmp = nullptr;
  size_t out_cap = 0, tmp_bytes = 0;
  int64_t total = 0;
  bool ok = true;
  auto CK = [&](cudaError_t e) { if (e != cudaSuccess) { ok = false; ctx->err = std::string("cmx_format_bed_gpu: ") + cudaGetErrorString(e); } return ok; };
  cudaStream_t st = ctx->stream;
  do {
    if (!CK(cudaMalloc(&d_names, cat.size() + 1)) || !CK(cudaMalloc(&d_noff, (n_seq + 1) * 4))) break;
    if (!CK(cudaMemcpyAsync(d_names, cat.data(), cat.size(), cudaMemcpyHostToDevice, st)) || !CK(cudaMemcpyAsync(d_noff, noff.data(), (n_seq + 1) * 4, cudaMemcpyHostToDevice, st))) break;
    if (nc == 0) break;
    if (!CK(cudaMalloc(&d_rec, nc * sizeof(PpRecord))) || !CK(cudaMalloc(&d_len, (nc + 1) * 4)) || !CK(cudaMalloc(&d_off, (nc + 1) * 8))) break;
    if (barcode_keys && !CK(cudaMalloc(&d_bc, nc * 8))) break;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_len, d_off, (int)nc + 1, st);
    if (!CK(cudaMalloc(&d_tmp, tmp_bytes))) break;
    for (u64 c0 = 0; c0 < n && ok; c0 += CH) {
      const u64 m = std::min(CH, n - c0);
      const unsigned nb = (unsigned)((m + 255) / 256);
      if (!CK(cudaMemcpyAsync(d_rec, records + c0, m * sizeof(PpRecord), cudaMemcpyHostToDevice, st))) break;
      if (barcode_keys && !CK(cudaMemcpyAsync(d_bc, barcode_keys + c0, m * 8, cudaMemcpyHostToDevice, st))) break;
      if (!CK(cudaMemsetAsync(d_len + m, 0, 4, st))) break;
      bed_len_kernel<<<nb, 256, 0, st>>>(d_rec, m, d_noff, barcode_keys ? (int)bc_len : 0, d_len);
      if (!CK(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_len, d_off, (int)m + 1, st))) break;
      u64 bytes = 0;
      if (!CK(cudaMemcpyAsync(&bytes, d_off + m, 8, cudaMemcpyDeviceToHost, st)) || !CK(cudaStreamSynchronize(st))) break;
      if (buf && total + (int64_t)bytes <= cap) {
        if (bytes > out_cap) { cudaFree(d_out); d_out = nullptr; out_cap = bytes + bytes / 8; if (!CK(cudaMalloc(&d_out, out_cap))) break; }
        bed_write_kernel<<<nb, 256, 0, st>>>(d_rec, d_bc, m, d_names, d_noff, barcode_keys ? (int)bc_len : 0, d_off, 0, d_out);
        if (!CK(cudaMemcpyAsync(buf + total, d_out, bytes, cudaMemcpyDeviceToHost, st)) || !CK(cudaStreamSynchronize(st))) break;
      }
      total += (int64_t)bytes;
    }
  } while (0);
  if (ok) CK(cudaGetLastError());
  cudaFree(d_names); cudaFree(d_noff); cudaFree(d_rec); cudaFree(d_len); cudaFree(d_off); cudaFree(d_bc); cudaFree(d_tmp); cudaFree(d_out);
  return ok ? total : -1;
}

// Pairs text on the device (pairs_len_kernel / pairs_write_kernel), byte-identical to cmx_format_pairs: the header is
// written by the host, the lines by one thread each.  Read names travel as one concatenation + offsets.
int64_t cmx_format_pairs_gpu(cmx_ctx *ctx, const char *const *names, const uint32_t *lengths, uint32_t n_seq, const cmx_pairs_record *records, uint64_t n,
                             const char *const *read_names, uint64_t n_read_names, uint32_t first_read_id, char *buf, int64_t cap) {
  if (!ctx || !names || !lengths || (!records && n) || (!read_names && n)) return -1;
  if (cudaSetDevice(ctx->device) != cudaSuccess) return -1;
  std::string hdr = "## pairs format v1.0.0\n#shape: upper triangle\n";  // mapping_writer.cc:383-402
  for (uint32_t i = 0; i < n_seq; ++i) hdr += std::string("#chromsize: ") + names[i] + " " + std::to_string(lengths[i]) + "\n";
  hdr += "#columns: readID chrom1 pos1 chrom2 pos2 strand1 strand2 pair_type mapq1 mapq2\n";
  int64_t total = (int64_t)hdr.size();
  if (buf && total <= cap) memcpy(buf, hdr.data(), hdr.size());
  if (n == 0) return total;
  std::string cat;
  std::vector<u32> noff(n_seq + 1, 0);
  for (u32 i = 0; i < n_seq; ++i) { cat += names[i]; noff[i + 1] = (u32)cat.size(); }
  std::vector<u64> roff(n_read_names + 1, 0);
  for (u64 i = 0; i < n_read_names; ++i) roff[i + 1] = roff[i] + strlen(read_names[i]);
  std::string rcat;
  rcat.resize(roff[n_read_names]);
  for (u64 i = 0; i < n_read_names; ++i) memcpy(&rcat[roff[i]], read_names[i], roff[i + 1] - roff[i]);
  char *d_names = nullptr, *d_rn = nullptr, *d_out = nullptr;
  u32 *d_noff = nullptr, *d_len = nullptr;
  u64 *d_roff = nullptr, *d_off = nullptr;
  PpRecord *d_rec = nullptr;
  void *d_tmp = nullptr;
  size_t tmp_bytes = 0;
  bool ok = true;
  auto CK = [&](cudaError_t e) { if (e != cudaSuccess) { ok = false; ctx->err = std::string("cmx_format_pairs_gpu: ") + cudaGetErrorString(e); } return ok; };
  cudaStream_t st = ctx->stream;
  const u64 CH = 8u << 20;
  const u64 nc = std::min<u64>(n, CH);
  do {
    if (!CK(cudaMalloc(&d_names, cat.size() + 1)) || !CK(cudaMalloc(&d_noff, (n_seq + 1) * 4)) || !CK(cudaMalloc(&d_rn, rcat.size() + 1)) ||
        !CK(cudaMalloc(&d_roff, (n_read_names + 1) * 8)) || !CK(cudaMalloc(&d_rec, nc * sizeof(PpRecord))) || !CK(cudaMalloc(&d_len, (nc + 1) * 4)) ||
        !CK(cudaMalloc(&d_off, (nc + 1) * 8)))
      break;
    if (!CK(cudaMemcpyAsync(d_names, cat.data(), cat.size(), cudaMemcpyHostToDevice, st)) || !CK(cudaMemcpyAsync(d_noff, noff.data(), (n_seq + 1) * 4, cudaMemcpyHostToDevice, st)) ||
        !CK(cudaMemcpyAsync(d_rn, rcat.data(), rcat.size(), cudaMemcpyHostToDevice, st)) || !CK(cudaMemcpyAsync(d_roff, roff.data(), (n_read_names + 1) * 8, cudaMemcpyHostToDevice, st)))
      break;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_len, d_off, (int)nc + 1, st);
    if (!CK(cudaMalloc(&d_tmp, tmp_bytes))) break;
    size_t out_cap = 0;
    for (u64 c0 = 0; c0 < n && ok; c0 += CH) {
      const u64 m = std::min(CH, n - c0);
      const unsigned nb = (unsigned)((m + 255) / 256);
      if (!CK(cudaMemcpyAsync(d_rec, records + c0, m * sizeof(PpRecord), cudaMemcpyHostToDevice, st)) || !CK(cudaMemsetAsync(d_len + m, 0, 4, st))) break;
      pairs_len_kernel<<<nb, 256, 0, st>>>(d_rec, m, d_noff, d_roff, first_read_id, d_len);
      if (!CK(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_len, d_off, (int)m + 1, st))) break;
      u64 bytes = 0;
      if (!CK(cudaMemcpyAsync(&bytes, d_off + m, 8, cudaMemcpyDeviceToHost, st)) || !CK(cudaStreamSynchronize(st))) break;
      if (buf && total + (int64_t)bytes <= cap) {
        if (bytes > out_cap) { cudaFree(d_out); d_out = nullptr; out_cap = bytes + bytes / 8; if (!CK(cudaMalloc(&d_out, out_cap))) break; }
        pairs_write_kernel<<<nb, 256, 0, st>>>(d_rec, m, d_names, d_noff, d_rn, d_roff, first_read_id, d_off, d_out);
        if (!CK(cudaMemcpyAsync(buf + total, d_out, bytes, cudaMemcpyDeviceToHost, st)) || !CK(cudaStreamSynchronize(st))) break;
      }
      total += (int64_t)bytes;
    }
  } while (0);
  if (ok) CK(cudaGetLastError());
  cudaFree(d_names); cudaFree(d_noff); cudaFree(d_rn); cudaFree(d_roff); cudaFree(d_rec); cudaFree(d_len); cudaFree(d_off); cudaFree(d_tmp); cudaFree(d_out);
  return ok ? total : -1;
}

// --TagAlign for paired-end records (mapping_writer.cc:84-110): one line per mate, the duplicate count on the second.
// (Single-end TagAlign lines are the BED lines, mapping_writer.cc:55-62.)
int64_t cmx_format_tagalign(const char *const *names, const cmx_pe_record *recs, uint64_t n, char *buf, int64_t cap) {
  int64_t len = 0;
  char line[2200];
  for (uint64_t i = 0; i < n; ++i) {
    const cmx_pe_record &r = recs[i];
    const uint32_t pos_end = r.fragment_start + r.positive_alignment_length, neg_end = r.fragment_start + r.fragment_length;
    const uint32_t neg_start = neg_end - r.negative_alignment_length;
    const char *nm = names[r.rid];
    int l;
    if (r.direction)
      l = snprintf(line, sizeof(line), "%s\t%u\t%u\tN\t%u\t+\n%s\t%u\t%u\tN\t%u\t-\t%u\n", nm, r.fragment_start, pos_end, (uint32_t)r.mapq, nm, neg_start, neg_end,
                   (uint32_t)r.mapq, (uint32_t)r.num_dups);
    else
      l = snprintf(line, sizeof(line), "%s\t%u\t%u\tN\t%u\t-\n%s\t%u\t%u\tN\t%u\t+\t%u\n", nm, neg_start, neg_end, (uint32_t)r.mapq, nm, r.fragment_start, pos_end,
                   (uint32_t)r.mapq, (uint32_t)r.num_dups);
    if (buf && len + l <= cap) memcpy(buf + len, line, l);
    len += l;
  }
  return len;
}

int64_t cmx_format_bed(const char *const *names, const cmx_pe_record *recs, uint64_t n, char *buf, int64_t cap) {
  int64_t len = 0;
  char line[1100];
  for (uint64_t i = 0; i < n; ++i) {  // mapping_writer.cc:75-83
    const cmx_pe_record &r = recs[i];
    const int l = snprintf(line, sizeof(line), "%s\t%u\t%u\tN\t%u\t%c\t%u\n", names[r.rid], r.fragment_start,
                           (uint32_t)(r.fragment_start + r.fragment_length), (uint32_t)r.mapq, r.direction ? '+' : '-', (uint32_t)r.num_dups);
    if (buf && len + l <= cap) memcpy(buf + len, line, l);
    len += l;
  }
  return len;
}

}  // extern "C"
