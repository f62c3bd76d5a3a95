// chromap_b200 — the one exchange step of the multi-GPU path: duplicate removal over the whole run,
// one process per GPU.  The reference is single-process; the semantics kept are those of its low-memory merge
// (mapping_writer.h:166-376): records in the reference's order (bed_mapping.h:145-159, 208-219), duplicates = equal
// (rid, fragment start, fragment length [, barcode]) (bed_mapping.h:154-159, 216-219), the first record in that order
// carrying the group's highest MAPQ survives (mapping_writer.h:268-270), its duplicate count saturates at 255 (:282-284),
// MAPQ filter after duplicate removal (:281).
//
// Every rank packs its records into compact tuples, ONE all-gather moves them over NVLink (ncclAllGather on the
// context's stream; an 8-byte all-gather of the counts sizes it), then every rank sorts the gathered tuples and decides
// which of ITS OWN records survive and with which duplicate count.  No record leaves its rank.
#pragma once
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>

#include "postprocess.cuh"

// 16 bytes (bulk) / 24 bytes (barcoded):  a = rid << 32 | fragment_start;  b = length << 48 | mapq << 40 | direction << 36 |
// is_unique << 32 | read_id (pp_key_word's second word);  bc = barcode key.  Unsigned order of (a, length, [bc,] low 48 bits of b)
// = the reference's record order.  Padding entries (ranks hold different numbers of records) carry a = ~0 and sort last.
#define EX_PAD 0xFFFFFFFFFFFFFFFFull

__global__ void ex_pack_kernel(const PpRecord *recs, const u64 *bcs, u64 n, u64 n_pad, int with_bc, u64 *tuples) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  const int tw = with_bc ? 3 : 2;
  u64 a = EX_PAD, b = EX_PAD, c = EX_PAD;
  if (i < n) {
    const PpRecord r = recs[i];
    a = ((u64)r.w[1] << 32) | r.w[2];
    b = ((u64)pe_len(r) << 48) | ((u64)pe_mapq(r) << 40) | ((u64)pe_dir(r) << 36) | ((u64)pe_uniq(r) << 32) | r.w[0];
    c = with_bc ? bcs[i] : 0ull;
  }
  tuples[i * tw] = a; tuples[i * tw + 1] = b;
  if (with_bc) tuples[i * tw + 2] = c;
}
// key word of the tuple at position idx[i]: pass 0 = low 48 bits of b, 1 = barcode, 2 = length, 3 = a, 4 = all of b
__global__ void ex_key_kernel(const u64 *tuples, int tw, int pass, const u32 *idx, u64 n, u64 *keys) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u64 *t = tuples + (u64)idx[i] * tw;
  keys[i] = pass == 0 ? (t[1] & 0x0000FFFFFFFFFFFFull) : pass == 1 ? t[2] : pass == 2 ? (t[1] >> 48) : pass == 4 ? t[1] : t[0];
}
// group heads in sorted order; head of a group = first record of a run of equal (a, length [, bc])
__global__ void ex_head_kernel(const u64 *tuples, int tw, int dedup, const u32 *idx, u64 n, u8 *head) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool h = true;
  if (dedup && i > 0) {
    const u64 *p = tuples + (u64)idx[i - 1] * tw, *q = tuples + (u64)idx[i] * tw;
    h = p[0] != q[0] || (p[1] >> 48) != (q[1] >> 48) || (tw == 3 && p[2] != q[2]);
  }
  head[i] = h ? 1 : 0;
}
// one thread per group head: survivor = first record of the run with the highest MAPQ; if it belongs to this rank and
// passes the MAPQ filter, mark it (keep[i] = 1, sel[i] = its local index, dups[i] = saturated run length)
__global__ void ex_resolve_kernel(const u64 *tuples, int tw, const u32 *idx, const u8 *head, u64 n, u64 n_pad, int rank, int mapq_threshold, u8 *keep,
                                  u32 *sel, u8 *dups) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  keep[i] = 0;
  if (!head[i]) return;
  u64 best = i;
  u32 best_q = (u32)((tuples[(u64)idx[i] * tw + 1] >> 40) & 0xFFu), run = 1;
  for (u64 j = i + 1; j < n && !head[j]; ++j) {
    const u32 q = (u32)((tuples[(u64)idx[j] * tw + 1] >> 40) & 0xFFu);
    if (q > best_q) { best_q = q; best = j; }
    ++run;
  }
  const u64 pos = idx[best];  // position in the gathered array = owner * n_pad + local index
  if ((int)best_q >= mapq_threshold && (int)(pos / n_pad) == rank) {
    keep[i] = 1;
    sel[i] = (u32)(pos % n_pad);
    dups[i] = (u8)(run > 255u ? 255u : run);
  }
}
__global__ void ex_gather_kernel(const PpRecord *recs, const u64 *bcs, const u32 *sel, const u8 *dups, int set_dups, u64 n, PpRecord *out, u64 *out_bc) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  PpRecord r = recs[sel[i]];
  if (set_dups) r.w[4] = (r.w[4] & 0xFFFF00FFu) | ((u32)dups[i] << 8);
  out[i] = r;
  if (bcs) out_bc[i] = bcs[sel[i]];
}

// ---- the same step as a range shuffle (cmx_dedup_shuffle) ---------------------------------------------------------------
// The all-gather above makes every rank sort the tuples of ALL ranks: its cost per rank grows with the size of the run.
// The shuffle keeps it at the size of the rank's share: the records' primary key word a = rid << 32 | fragment_start is
// range-partitioned over the ranks (splitters from an all-gathered sample — a sample sort), each record travels once, to
// the rank that owns its key range (grouped ncclSend / ncclRecv over NVLink), and that rank runs the ordinary
// single-GPU post-processing (postprocess.cuh: the reference's order, duplicate rule, MAPQ filter, Tn5) on what it
// received.  Duplicates agree in `a`, so a group never straddles two ranks; the run's output is the ranks' outputs in
// rank order.
#define SH_SAMPLE 4096  // sample keys per rank

__global__ void sh_key_kernel(const PpRecord *recs, u64 n, u64 *a) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = ((u64)recs[i].w[1] << 32) | recs[i].w[2];
}
// SH_SAMPLE evenly spaced records of this rank (the local order is the mapping order: unrelated to the key); ranks with
// fewer records than that pad with EX_PAD, which the splitter choice ignores
__global__ void sh_sample_kernel(const u64 *a, u64 n, u64 *sample) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= SH_SAMPLE) return;
  const u64 take = n < SH_SAMPLE ? n : SH_SAMPLE;
  sample[i] = i < take ? a[(u64)i * n / take] : EX_PAD;
}
// destination rank = number of splitters <= a  (rank r owns  splitter[r-1] <= a < splitter[r])
__global__ void sh_dest_kernel(const u64 *a, u64 n, const u64 *splitters, int n_split, u32 *dest) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u64 x = a[i];
  int lo = 0, hi = n_split;
  while (lo < hi) { const int m = (lo + hi) >> 1; if (splitters[m] <= x) lo = m + 1; else hi = m; }
  dest[i] = (u32)lo;
}
// first position of every destination in the destination-sorted order: off[d] = lower_bound(sorted, d), off[R] = n
__global__ void sh_bounds_kernel(const u32 *sorted_dest, u64 n, int R, u64 *off) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d > R) return;
  u64 lo = 0, hi = n;
  while (lo < hi) { const u64 m = (lo + hi) >> 1; if (sorted_dest[m] < (u32)d) lo = m + 1; else hi = m; }
  off[d] = lo;
}
