// chromap-b200 — command-line front end that keeps the reference's CLI for the supported path
// (chromap_driver.cc:16-159 option names, :247-275 presets, :451-531 checks): index construction (-i) on the
// GPU with the reference's index file format, and paired-end mapping to BED through the C ABI.
// Everything here is host plumbing around cmx_map_batch_pe; the batch loop replaces chromap.h:851-1290.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <algorithm>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../../include/chromap_b200.h"
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "seqio.h"

using cmxhost::IndexFile;
using cmxhost::IndexMap;
using cmxhost::Reference;
using cmxhost::SeqReader;

static void Die(const std::string &msg) {  // utils.h:71-74
  fprintf(stderr, "%s\n", msg.c_str());
  exit(255);
}
static double Now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Batch {
  std::string s1, s2;
  std::vector<uint32_t> o1{0}, o2{0};
  std::vector<std::string> names1;  // read-1 names, kept for pairs / SAM output only
  std::vector<std::string> names2;  // SAM only
  std::string q1, q2;               // qualities in the layout of s1 / s2 (SAM only)
  std::string bc, bq;               // cell barcodes + qualities, bc_len bytes per pair (scATAC)
  uint32_t n = 0, first_id = 0;
  bool dev = false;      // reads were packed on the device (cmx_ingest_fastq): dev_in holds device pointers
  cmx_batch dev_in{};
  void Clear() { dev = false; s1.clear(); s2.clear(); o1.assign(1, 0); o2.assign(1, 0); names1.clear(); names2.clear(); q1.clear(); q2.clear(); bc.clear(); bq.clear(); n = 0; }
};

// Raw text of one read file for the device-side FASTQ parser.  Plain files are read with read(2) straight into a page-locked
// buffer (no zlib copy, H2D at PCIe speed); gzip files go through gzread.  Whole 4-line records are cut off the front of the
// buffer (same rule as cmx_fastq_cut, but the scan remembers where it stopped instead of starting over after every refill),
// the rest stays for the next batch.
struct RawBuf {  // uninitialised page-aligned bytes (a std::vector would zero-fill hundreds of megabytes before they are read into)
  char *p = nullptr;
  size_t n = 0;
  char *data() { return p; }
  const char *data() const { return p; }
  size_t size() const { return n; }
  char &operator[](size_t i) { return p[i]; }
  void Grow(size_t bytes, size_t keep) {
    char *q = (char *)aligned_alloc(4096, (bytes + 4095) & ~(size_t)4095);
    if (!q) { fprintf(stderr, "chromap-b200: out of memory\n"); exit(255); }
    if (keep) memcpy(q, p, keep);
    free(p);
    p = q; n = bytes;
  }
  ~RawBuf() { free(p); }
};
struct RawFile {
  gzFile f = nullptr;
  int fd = -1;
  RawBuf buf;
  size_t have = 0;
  bool eof = false, pinned = false;
  size_t fpos = 0;                   // plain files: offset of the next unread byte
  size_t scan = 0;                   // scan state: bytes examined (see Fill)
  size_t chunk = 0;                  // bytes per read; 0 = 32 MB (gzip) / 128 MB (plain).  Set by the host test only.
  size_t file_bytes = 0, first_cut = 0;  // plain files: size; bytes of the first batch handed out (to size the record store)
  bool Open(const std::string &path) {
    Close();
    have = 0; eof = false; scan = 0; nl = 0; fpos = 0;
    unsigned char magic[2] = {0, 0};
    FILE *t = fopen(path.c_str(), "rb");
    if (!t) return false;
    const size_t got = fread(magic, 1, 2, t);
    fclose(t);
    if (got == 2 && magic[0] == 0x1f && magic[1] == 0x8b) { f = gzopen(path.c_str(), "rb"); if (f) gzbuffer(f, 1 << 20); return f != nullptr; }
    fd = open(path.c_str(), O_RDONLY);
    if (fd >= 0) { posix_fadvise(fd, 0, 0, POSIX_FADV_SEQUENTIAL); struct stat sb; if (fstat(fd, &sb) == 0) file_bytes = (size_t)sb.st_size; }
    return fd >= 0;
  }
  void Close() {
    if (f) gzclose(f);
    if (fd >= 0) close(fd);
    f = nullptr; fd = -1;
    if (pinned) { cmx_host_unregister(buf.data()); pinned = false; }
  }
  void Reserve(size_t bytes, bool exact = false) {  // grow (rarely): the buffer is pinned once it has its working size
    if (buf.size() >= bytes) return;
    if (pinned) { cmx_host_unregister(buf.data()); pinned = false; }
    buf.Grow(exact ? bytes : bytes + bytes / 4, have);
    pinned = cmx_host_register(buf.data(), buf.size()) == 0;
  }
  // Size (and page-lock) the buffer for calls of max_records before the first one: bytes per record from the head of the file.
  // Run at start-up, beside the index upload, so that the mapping phase never allocates.
  void Prepare(uint32_t max_records) {
    if (f || fd < 0 || !file_bytes) return;
    std::vector<char> head(1u << 20);
    const ssize_t got = pread(fd, head.data(), head.size(), 0);
    if (got <= 0) return;
    const uint64_t lines = CountNewlines(head.data(), (size_t)got);
    if (lines < 8) return;
    const double per_record = (double)got / ((double)lines / 4.0);
    const size_t want = (size_t)(per_record * max_records * 1.1) + (128u << 20) + 4096;
    Reserve(std::min(want, file_bytes + 4096), true);
  }
  // bytes of up to max_records whole records now in the buffer (reading more as needed); *n = their number.
  // Newlines are counted in bulk (SSE2 compare + popcount, by the threads that read the data); only the stretch that holds
  // the last newline wanted is walked line by line.  State: nl = newlines in buf[0, scan).
  uint64_t nl = 0;
  static uint64_t CountNewlines(const char *p, size_t n) {
    uint64_t c = 0;
    size_t i = 0;
#if defined(__SSE2__)
    const __m128i v = _mm_set1_epi8('\n');
    for (; i + 64 <= n; i += 64) {
      const unsigned m0 = (unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(_mm_loadu_si128((const __m128i *)(p + i)), v));
      const unsigned m1 = (unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(_mm_loadu_si128((const __m128i *)(p + i + 16)), v));
      const unsigned m2 = (unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(_mm_loadu_si128((const __m128i *)(p + i + 32)), v));
      const unsigned m3 = (unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(_mm_loadu_si128((const __m128i *)(p + i + 48)), v));
      c += (uint64_t)__builtin_popcountll(((uint64_t)m0) | ((uint64_t)m1 << 16) | ((uint64_t)m2 << 32) | ((uint64_t)m3 << 48));
    }
#endif
    for (; i < n; ++i) c += p[i] == '\n';
    return c;
  }
  // advance the scan over buf[scan, scan + len) whose newline count is known; returns true once 4 * max_records are in
  bool Advance(size_t len, uint64_t count, uint64_t target) {
    if (nl + count < target) { nl += count; scan += len; return false; }
    while (nl < target) {  // the last newline wanted lies in this stretch
      const void *q = memchr(buf.data() + scan, '\n', have - scan);
      scan = (size_t)((const char *)q - buf.data()) + 1;
      ++nl;
    }
    return true;
  }
  uint64_t Fill(uint32_t max_records, uint32_t *n) {
    const uint64_t target = 4ull * max_records;
    if (max_records == 0) { *n = 0; return 0; }
    if (scan < have && nl < target) {  // what an earlier call left behind (less than one read chunk)
      const size_t len = have - scan;
      if (Advance(len, CountNewlines(buf.data() + scan, len), target)) { *n = max_records; return scan; }
    }
    for (;;) {
      if (nl >= target) { *n = max_records; return scan; }
      if (eof) {  // the file ended first: whole records only; up to three lines of an unfinished one stay behind
        const uint64_t whole = nl / 4;
        size_t e = have;
        for (uint64_t extra = nl - 4 * whole; extra > 0 && e > 0; --extra) {
          const void *q = memrchr(buf.data(), '\n', e - 1);  // the newline before the last line's own
          e = q ? (size_t)((const char *)q - buf.data()) + 1 : 0;
        }
        if (nl == 0) e = 0;
        *n = (uint32_t)whole;
        return whole ? e : 0;
      }
      size_t want = chunk ? chunk : (f ? (32u << 20) : (128u << 20));
      if (!f && !chunk && file_bytes)  // a plain file: never ask for more than is left (+ a few bytes, so that the end is seen as a short read)
        want = std::max<size_t>(64, std::min(want, file_bytes - std::min(file_bytes, fpos) + 8));
      Reserve(have + want + 1);
      if (f) {
        const long got = gzread(f, buf.data() + have, (unsigned)want);
        if (got <= 0) eof = true;
        else {
          have += (size_t)got;
          if (Advance((size_t)got, CountNewlines(buf.data() + scan, (size_t)got), target)) { *n = max_records; return scan; }
        }
      } else {  // plain file: eight threads pread() a slice each (one thread copies out of the page cache at 3-5 GB/s)
        const int nt = 8;
        const size_t slice = want / nt;
        size_t part[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        uint64_t cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        std::thread th[8];
        for (int t = 0; t < nt; ++t)
          th[t] = std::thread([&, t]() {
            size_t done = 0;
            char *dst = buf.data() + have + t * slice;
            while (done < slice) {
              const ssize_t r = pread(fd, dst + done, slice - done, (off_t)(fpos + t * slice + done));
              if (r <= 0) break;
              done += (size_t)r;
            }
            part[t] = done;
            cnt[t] = CountNewlines(dst, done);
          });
        for (int t = 0; t < nt; ++t) th[t].join();
        size_t got = 0;
        int used = 0;
        for (int t = 0; t < nt; ++t) { got += part[t]; ++used; if (part[t] < slice) break; }  // a short slice is the end of the file
        fpos += got;
        if (got == 0) eof = true;
        else {
          have += got;
          bool done = false;
          for (int t = 0; t < used && !done; ++t) done = Advance(part[t], cnt[t], target);
          if (done) { *n = max_records; return scan; }
        }
      }
      if (eof && have > 0 && buf[have - 1] != '\n') { buf[have++] = '\n'; ++nl; scan = have; }  // a last line without its newline
    }
  }
  void Consume(uint64_t bytes) {  // the remainder (less than one read chunk) is counted again by the next Fill
    if (!first_cut) first_cut = bytes;
    memmove(buf.data(), buf.data() + bytes, have - bytes);
    have -= bytes;
    scan = 0; nl = 0;
  }
};

static double g_t_fill = 0, g_t_ingest = 0;  // loader thread: reading + cutting, device-side parsing (summed over calls)

// One batch through the device-side parser.  Returns false if a file is not plain 4-line FASTQ (the caller then uses the
// host reader); dies on real input errors, like LoadBatch.
static bool LoadBatchGpu(cmx_ctx *ctx, RawFile *f1, RawFile *f2, RawFile *fb, int parity, uint32_t max_pairs, Batch *b, bool keep_names, uint32_t bc_len) {
  b->Clear();
  uint32_t n1 = 0, n2 = 0, nb = 0;
  uint64_t c1 = 0, c2 = 0, cb = 0;  // the files are read (and inflated) side by side
  const double t_f0 = Now();
  std::thread t2, tb;
  if (f2) t2 = std::thread([&]() { c2 = f2->Fill(max_pairs, &n2); });
  if (fb) tb = std::thread([&]() { cb = fb->Fill(max_pairs, &nb); });
  c1 = f1->Fill(max_pairs, &n1);
  if (f2) t2.join();
  if (fb) tb.join();
  const double t_f1 = Now();
  g_t_fill += t_f1 - t_f0;
  if ((f2 && n2 != n1) || (fb && nb != n1)) Die("Numbers of reads and barcodes don't match!");
  if (n1 == 0) {
    if (f1->have || (f2 && f2->have) || (fb && fb->have)) return false;  // trailing bytes that are no whole record: not 4-line FASTQ
    return true;
  }
  cmx_ingested g1{}, g2{}, gb{};
  std::vector<uint32_t> spans;
  if (keep_names) spans.resize(2 * (size_t)n1);
  if (cmx_ingest_fastq(ctx, parity * 3 + 0, f1->buf.data(), c1, 0, keep_names ? spans.data() : nullptr, &g1)) return false;
  if (f2 && cmx_ingest_fastq(ctx, parity * 3 + 1, f2->buf.data(), c2, 0, nullptr, &g2)) return false;
  if (fb) {
    if (cmx_ingest_fastq(ctx, parity * 3 + 2, fb->buf.data(), cb, 1, nullptr, &gb)) return false;
    if (gb.min_len != bc_len || gb.max_len != bc_len) Die("ERROR: barcode lengths are not equal in the sample!");
  }
  if (keep_names) for (uint32_t i = 0; i < n1; ++i) b->names1.emplace_back(f1->buf.data() + spans[2 * i], spans[2 * i + 1]);
  g_t_ingest += Now() - t_f1;
  const double t_c0 = Now();
  f1->Consume(c1);
  if (f2) f2->Consume(c2);
  if (fb) fb->Consume(cb);
  g_t_fill += Now() - t_c0;
  b->n = n1; b->dev = true;
  b->dev_in.n_pairs = n1; b->dev_in.seq1 = g1.seq; b->dev_in.off1 = g1.off; b->dev_in.on_device = 1;
  if (f2) { b->dev_in.seq2 = g2.seq; b->dev_in.off2 = g2.off; }
  if (fb) { b->dev_in.bc_seq = gb.seq; b->dev_in.bc_qual = gb.qual; b->dev_in.bc_len = bc_len; }
  return true;
}

// LoadPairedEndReadsWithBarcodes (chromap.cc:93-174, non-barcode): empty reads are skipped per file
// (sequence_batch.cc:28-31), the two files must run out together.
// utils.h:107-126
static uint64_t BarcodeSeed(const std::string &s) {
  uint64_t seed = 0;
  for (char c : s) {
    const int b = (c == 'A' || c == 'a') ? 0 : (c == 'C' || c == 'c') ? 1 : (c == 'G' || c == 'g') ? 2 : (c == 'T' || c == 't') ? 3 : 4;
    seed = b < 4 ? (seed << 2) | (uint64_t)b : seed << 2;
  }
  return seed;
}

static uint32_t LoadBatch(SeqReader &r1, SeqReader &r2, uint32_t max_pairs, Batch *b, bool keep_names, SeqReader *rb = nullptr, uint32_t bc_len = 0,
                          bool se = false, bool keep_sam = false) {
  std::string n, s, q;
  b->Clear();
  while (b->n < max_pairs) {
    bool a = r1.Next(&n, &s, &q);
    while (a && s.empty()) a = r1.Next(&n, &s, &q);
    if (a) { b->s1 += s; b->o1.push_back((uint32_t)b->s1.size()); if (keep_names) b->names1.push_back(n); if (keep_sam) { q.resize(s.size(), 'I'); b->q1 += q; } }
    bool c = a;  // single-end: no second file
    if (!se) {
      c = r2.Next(&n, &s, &q);
      while (c && s.empty()) c = r2.Next(&n, &s, &q);
      if (c) { b->s2 += s; b->o2.push_back((uint32_t)b->s2.size()); if (keep_sam) { b->names2.push_back(n); q.resize(s.size(), 'I'); b->q2 += q; } }
    }
    bool d = c;
    if (rb) {
      d = rb->Next(&n, &s, &q);
      while (d && s.empty()) d = rb->Next(&n, &s, &q);
      if (d) {
        if (s.size() != bc_len) Die("ERROR: barcode lengths are not equal in the sample!");
        q.resize(bc_len, 'I');
        b->bc += s; b->bq += q;
      }
    }
    if (r1.Corrupted() || (!se && r2.Corrupted()) || (rb && rb->Corrupted())) Die("Didn't reach the end of sequence file, which might be corrupted!");  // sequence_batch.cc:46-55
    if (!a && !c && !d) break;
    if (a != c || c != d) Die("Numbers of reads and barcodes don't match!");
    ++b->n;
  }
  return b->n;
}

int main(int argc, char **argv) {
  cmx_params p;
  cmx_default_params(&p);
  std::string preset, ref_path, index_path, r1_path, r2_path, out_path, bc_path, wl_path;
  int bc_err = 1, out_nw = 0;
  bool skip_bc_check = false, host_reader = false, paf = false;
  bool cell_level_dedup = false;  // remove_pcr_duplicates_at_bulk_level == false (mapping_parameters.h:49; --preset atac clears it)
  double bc_prob = 0.9;
  bool build_index = false, bed = false, user_set_format = false;
  int k = 17, w = 7, threads = 1;
  (void)threads;
  for (int i = 1; i < argc; ++i)
    if (!strcmp(argv[i], "--preset") && i + 1 < argc) preset = argv[i + 1];
  if (cmx_apply_preset(&p, preset.c_str()) != 0) Die("Unrecognized preset parameters " + preset + "\n");
  if (preset == "atac") cell_level_dedup = true;  // chromap_driver.cc:254
  if (!preset.empty()) fprintf(stderr, "Preset parameters for %s are used.\n", preset.c_str());
  for (int i = 1; i + 1 < argc; ++i)
    if (!strcmp(argv[i], "--min-frag-length")) { const int l = atoi(argv[i + 1]); if (l <= 60) { k = 17; w = 7; } else if (l <= 80) { k = 19; w = 10; } else { k = 23; w = 11; } }
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto val = [&]() -> std::string { if (i + 1 >= argc) Die("Option " + a + " is missing an argument"); return argv[++i]; };
    if (a == "--preset") val();
    else if (a == "-i" || a == "--build-index") build_index = true;
    else if (a == "-h" || a == "--help") { printf("chromap-b200: chromap's paired-end BED path on B200 GPUs (subset of chromap options; see DESIGN.md)\n"); return 0; }
    else if (a == "-v" || a == "--version") { fprintf(stderr, "chromap-b200 0.1 (parity target: chromap 0.3.3-r521)\n"); return 0; }
    else if (a == "-r" || a == "--ref") ref_path = val();
    else if (a == "-x" || a == "--index") index_path = val();
    else if (a == "-1" || a == "--read1") r1_path = val();
    else if (a == "-2" || a == "--read2") r2_path = val();
    else if (a == "-o" || a == "--output") out_path = val();
    else if (a == "-t" || a == "--num-threads") threads = atoi(val().c_str());
    else if (a == "-k" || a == "--kmer") k = atoi(val().c_str());
    else if (a == "-w" || a == "--window") w = atoi(val().c_str());
    else if (a == "--min-frag-length") val();  // applied before this loop: -k / -w always win over it (chromap_driver.cc:277-295)
    else if (a == "-e" || a == "--error-threshold") p.error_threshold = atoi(val().c_str());
    else if (a == "-s" || a == "--min-num-seeds") p.min_num_seeds = atoi(val().c_str());
    else if (a == "-f" || a == "--max-seed-frequencies") { const std::string v = val(); if (sscanf(v.c_str(), "%d,%d", &p.max_seed_freq0, &p.max_seed_freq1) != 2) Die("-f expects two comma separated integers"); }
    else if (a == "-l" || a == "--max-insert-size") p.max_insert_size = atoi(val().c_str());
    else if (a == "-q" || a == "--MAPQ-threshold") p.mapq_threshold = atoi(val().c_str());
    else if (a == "--min-read-length") p.min_read_length = atoi(val().c_str());
    else if (a == "--trim-adapters") p.trim_adapters = 1;
    else if (a == "--remove-pcr-duplicates") p.remove_pcr_duplicates = 1;
    else if (a == "--Tn5-shift") p.tn5_shift = 1;
    else if (a == "--low-mem") p.low_memory_mode = 1;
    else if (a == "--BED") { bed = true; user_set_format = true; }
    else if (a == "--split-alignment") p.split_alignment = 1;
    else if (a == "--pairs") p.output_format = 5;
    else if (a == "-b" || a == "--barcode") bc_path = val();
    else if (a == "--barcode-whitelist") wl_path = val();
    else if (a == "--bc-error-threshold") bc_err = atoi(val().c_str());
    else if (a == "--bc-probability-threshold") bc_prob = atof(val().c_str());
    else if (a == "--output-mappings-not-in-whitelist") out_nw = 1;
    else if (a == "--skip-barcode-check") skip_bc_check = true;
    else if (a == "--host-reader") host_reader = true;  // parse FASTQ on the host (multi-line records, FASTA reads)
    else if (a == "-n" || a == "--max-num-best-mappings") p.max_num_best_mappings = atoi(val().c_str());
    else if (a == "--drop-repetitive-reads") p.drop_repetitive_reads = atoi(val().c_str());
    else if (a == "--remove-pcr-duplicates-at-cell-level") cell_level_dedup = true;   // chromap_driver.cc:395-400: the level only
    else if (a == "--remove-pcr-duplicates-at-bulk-level") cell_level_dedup = false;
    // options that cannot change BED / TagAlign / pairs output: the candidate cache (result-transparent), SAM scoring, QC estimators
    else if (a == "--cache-size" || a == "--cache-update-param" || a == "--frip-est-params" || a == "--k-for-minhash" || a == "-A" || a == "--match-score" ||
             a == "-B" || a == "--mismatch-penalty" || a == "-O" || a == "--gap-open-penalties" || a == "-E" || a == "--gap-extension-penalties") val();
    else if (a == "--debug-cache" || a == "--turn-off-num-uniq-cache-slots") {}
    else if (a == "--chr-order" || a == "--pairs-natural-chr-order" || a == "--read-format" || a == "--barcode-translate" || a == "--allocate-multi-mappings" ||
             a == "-p" || a == "--matrix-output-prefix")
      Die("chromap-b200: option " + a + " changes the output in ways that are not on the GPU path; use the reference chromap for it");
    else if (a == "--TagAlign") p.output_format = 2;  // same records as BED, TagAlign / PairedTagAlign text (chromap_driver.cc:417-418)
    else if (a == "--SAM") p.output_format = 4;  // device: ksw spans, CIGARs, MAPQ; host: flags, NM / MD, order, text (cmx_format_sam)
    else if (a == "--PAF") paf = true;  // BED-path records, PAF text and order on the host (cmx_format_paf)
    else if (a == "--summary")
      Die("chromap-b200: option " + a + " is not on the GPU path yet (BED and Hi-C pairs only); use the reference chromap for it");
    else Die("Unknown option " + a);
  }
  (void)bed; (void)user_set_format;
  if (!(((p.output_format == 1 || p.output_format == 2 || p.output_format == 4) && !p.split_alignment) || (p.output_format == 5 && p.split_alignment)))
    Die("chromap-b200: supported outputs are BED / TagAlign (no split alignment) and Hi-C pairs (--split-alignment --pairs / --preset hic)");
  const bool tagalign = p.output_format == 2;
  const bool sam = p.output_format == 4;
  const bool pairs = p.output_format == 5;
  cmx_ctx *ctx = nullptr;
  const double t_start = Now();
  if (build_index) {  // chromap_driver.cc:451-471
    if (ref_path.empty() || out_path.empty()) Die("No reference specified!");
    fprintf(stderr, "Build index for the reference.\nKmer length: %d, window size: %d\nReference file: %s\nOutput file: %s\n", k, w, ref_path.c_str(), out_path.c_str());
    Reference ref;
    if (!ref.Load(ref_path)) Die("Cannot find sequence file " + ref_path);
    int rc = cmx_create(&ctx, 0, &p);
    if (rc) Die(rc == CMX_ERR_NO_DEVICE ? "chromap-b200: no CUDA device (there is no CPU fallback)" : "chromap-b200: cmx_create failed");
    if (cmx_upload_reference(ctx, (uint32_t)ref.names.size(), ref.offsets.data(), ref.concat.data())) Die(cmx_last_error(ctx));
    if (cmx_build_index(ctx, k, w)) Die(cmx_last_error(ctx));
    IndexFile ix;
    ix.k = k; ix.w = w;
    uint32_t n_occ = 0;
    if (cmx_download_index(ctx, &ix.n_buckets, &ix.size, nullptr, nullptr, nullptr, &n_occ, nullptr)) Die(cmx_last_error(ctx));
    ix.flags.resize(ix.n_buckets < 16 ? 1 : ix.n_buckets >> 4); ix.keys.resize(ix.n_buckets); ix.vals.resize(ix.n_buckets); ix.occ.resize(n_occ);
    if (cmx_download_index(ctx, &ix.n_buckets, &ix.size, ix.flags.data(), ix.keys.data(), ix.vals.data(), &n_occ, ix.occ.data())) Die(cmx_last_error(ctx));
    ix.n_occupied = ix.size;
    ix.upper_bound = (uint32_t)(ix.n_buckets * 0.77 + 0.5);
    if (!ix.Save(out_path)) Die("Cannot write index file " + out_path);
    fprintf(stderr, "Lookup table size: %u, # buckets: %u, occurrence table size: %u.\nBuilt and saved index in %.2fs.\n", ix.size, ix.n_buckets, n_occ, Now() - t_start);
    cmx_destroy(ctx);
    return 0;
  }
  if (ref_path.empty()) Die("No reference specified!");
  if (index_path.empty()) Die("No index specified!");
  if (r1_path.empty()) Die("No read file specified!");
  const bool se = r2_path.empty();  // chromap_driver.cc:704-761: -1 alone = single-end
  if (se && pairs) Die("chromap-b200: pairs output needs paired-end reads");
  if ((tagalign || sam || paf) && !bc_path.empty()) Die("chromap-b200: --TagAlign / --SAM / --PAF with barcodes is not on the GPU path");
  if (paf && (sam || tagalign || pairs || p.trim_adapters)) Die("chromap-b200: --PAF goes with BED-path mapping without adapter trimming (trimmed read lengths are not returned yet)");
  if (pairs && p.remove_pcr_duplicates && !p.low_memory_mode)  // RemovePCRDuplicate keeps the LAST record of a run (mapping_processor.h:181-197): not on the GPU path for pairs
    Die("chromap-b200: duplicate removal of Hi-C pairs needs --low-mem (or --preset hic)");
  if (!bc_path.empty() && p.remove_pcr_duplicates && p.low_memory_mode && !cell_level_dedup)  // mapping_writer.h:254-262: only the low-memory merge has the bulk-level variant
    Die("chromap-b200: bulk-level duplicate removal of barcoded data is not on the GPU path (use --preset atac or --remove-pcr-duplicates-at-cell-level)");
  if (out_path.empty()) Die("No output file specified!");
  // start-up, three things at once (the reference does them one after the other, chromap.h:684-730): the reference sequences
  // are parsed by one thread, the index file is mapped and read ahead by the kernel, the CUDA context comes up on this thread
  Reference ref;
  bool ref_ok = false;
  double t_ref_parsed = 0;
  std::thread ref_loader([&]() { ref_ok = ref.Load(ref_path); t_ref_parsed = Now(); });
  IndexMap ix;
  const bool ix_ok = ix.Open(index_path);
  p.single_end = se ? 1 : 0;
  int rc = cmx_create(&ctx, 0, &p);
  if (!ix_ok) Die("Cannot load index file " + index_path);
  if (rc) Die(rc == CMX_ERR_NO_DEVICE ? "chromap-b200: no CUDA device (there is no CPU fallback)" : "chromap-b200: unsupported parameter combination");
  // the index arrays go to the device while the reference sequences are still being parsed — and while the read files are
  // opened, their page-locked buffers sized from the first records, and the first call's reads brought in
  RawFile g1, g2, gb;
  bool gpu_reader = !host_reader && !sam && !paf;  // SAM / PAF keep names (SAM: bases and qualities too) of every read on the host
  bool raw_ready = false;
  const uint32_t first_call_pairs = (uint32_t)p.batch_size * 4u;
  std::thread raw_prep;
  if (gpu_reader) {
    raw_prep = std::thread([&]() {
      const bool with_bc = !bc_path.empty();
      if (!g1.Open(r1_path) || (!se && !g2.Open(r2_path)) || (with_bc && !gb.Open(bc_path))) return;  // reported by open_all below
      auto warm = [&](RawFile *f) { uint32_t n = 0; f->Prepare(first_call_pairs); f->Fill(first_call_pairs, &n); };
      std::thread a, b;
      if (!se) a = std::thread([&]() { warm(&g2); });
      if (with_bc) b = std::thread([&]() { warm(&gb); });
      warm(&g1);
      if (a.joinable()) a.join();
      if (b.joinable()) b.join();
      raw_ready = true;
    });
  }
  const double t_ix = Now();
  if (cmx_upload_index(ctx, ix.k, ix.w, ix.n_buckets, ix.flags, ix.keys, ix.vals, ix.occ, ix.n_occ)) Die(cmx_last_error(ctx));
  fprintf(stderr, "Kmer size: %d, window size: %d.\nLookup table size: %u, occurrence table size: %u.\n", ix.k, ix.w, ix.size, ix.n_occ);
  ix.Close();
  const double t_ix_done = Now();
  ref_loader.join();
  if (!ref_ok) Die("Cannot find sequence file " + ref_path);
  fprintf(stderr, "Loaded all sequences successfully, number of sequences: %zu, number of bases: %zu.\n", ref.names.size(), ref.concat.size());
  const double t_ref = Now();
  if (cmx_upload_reference(ctx, (uint32_t)ref.names.size(), ref.offsets.data(), ref.concat.data())) Die(cmx_last_error(ctx));
  fprintf(stderr, "Start-up: context %.2fs, index to the device %.2fs, reference parsed (concurrently) after %.2fs, reference to the device %.2fs.\n", t_ix - t_start,
          t_ix_done - t_ix, t_ref_parsed - t_start, Now() - t_ref);
  if (raw_prep.joinable()) raw_prep.join();
  fprintf(stderr, "Reference and index resident on the device after %.2fs.\n", Now() - t_start);
  // scATAC pre-pass (chromap.h:755-761): barcode length from the first record, whitelist, abundance over the first >= 20 M
  // whitelisted barcodes (chromap.cc:364-386, 388-548)
  const bool sc = !bc_path.empty();
  uint32_t bc_len = 0;
  if (sc) {
    if (pairs) Die("chromap-b200: barcodes with Hi-C pairs output are not on the GPU path");
    SeqReader rb0;
    if (!rb0.Open(bc_path)) Die("Cannot find sequence file " + bc_path);
    std::string n, s, q;
    if (!rb0.Next(&n, &s, &q)) Die("Empty barcode file");
    bc_len = (uint32_t)s.size();
    if (bc_len > 32) Die("ERROR: barcode length is greater than 32!");
    if (!wl_path.empty()) {
      std::unordered_map<uint64_t, uint32_t> wl;
      gzFile f = gzopen(wl_path.c_str(), "r");
      if (!f) Die("ERROR: barcode whitelist file does not exist or is truncated!");
      char buf[256];
      while (gzgets(f, buf, sizeof(buf)) != NULL) {
        size_t l = strlen(buf);
        if (l && buf[l - 1] == '\n') buf[--l] = 0;
        if (l != bc_len) Die(wl.empty() ? "ERROR: whitelist and input barcode lengths are not equal!" : "ERROR: barcode lengths are not equal in the whitelist!");
        wl.emplace(BarcodeSeed(std::string(buf, l)), 0u);
      }
      gzclose(f);
      fprintf(stderr, "Loaded %zu barcodes.\n", wl.size());
      uint64_t num_sample = 0, in_batch = 0, loaded = 0;
      bool more = true;
      while (more) {  // one reference batch at a time; stop after the batch that reaches 20 M
        if (s.find('N') == std::string::npos) { auto it = wl.find(BarcodeSeed(s)); if (it != wl.end()) { ++it->second; ++num_sample; } }
        ++in_batch; ++loaded;
        more = rb0.Next(&n, &s, &q);
        while (more && s.empty()) more = rb0.Next(&n, &s, &q);
        if (in_batch == (uint64_t)p.batch_size || !more) {
          if (!skip_bc_check && num_sample * 20 < in_batch) Die("Less than 5% barcodes can be found or corrected based on the barcode whitelist.");
          if (num_sample >= 20000000ull) break;
          in_batch = 0;
        }
      }
      fprintf(stderr, "Compute barcode abundance using %llu.\n", (unsigned long long)num_sample);
      std::vector<uint64_t> keys; std::vector<uint32_t> counts;
      for (const auto &kv : wl) { keys.push_back(kv.first); counts.push_back(kv.second); }
      if (cmx_upload_barcode_whitelist(ctx, keys.data(), counts.data(), keys.size(), num_sample, bc_len, bc_err, bc_prob, out_nw)) Die(cmx_last_error(ctx));
    }
  }
  // double-buffered batch loop: the loader thread prepares batch b+1 while the GPU maps batch b (chromap.h:871-877).
  // Reads are parsed on the device (cmx_ingest_fastq) when the files are plain 4-line FASTQ, else by the host reader.
  SeqReader r1, r2, rb;
  Batch cur, next;
  auto open_all = [&](bool raw) {
    if (raw) {
      if (raw_ready) return;  // opened, sized and filled at start-up
      if (!g1.Open(r1_path)) Die("Cannot find sequence file " + r1_path);
      if (!se && !g2.Open(r2_path)) Die("Cannot find sequence file " + r2_path);
      if (sc && !gb.Open(bc_path)) Die("Cannot find sequence file " + bc_path);
    } else {
      if (sc && !rb.Open(bc_path)) Die("Cannot find sequence file " + bc_path);
      if (!r1.Open(r1_path)) Die("Cannot find sequence file " + r1_path);
      if (!se && !r2.Open(r2_path)) Die("Cannot find sequence file " + r2_path);
    }
  };
  int parity = 0;
  // one call carries four reference batches (chromap.h:182: 500 000 pairs each) when the reads are parsed on the device: the
  // library runs them as overlapping lanes; batch boundaries — which the multi-mapper sampling depends on — stay where they are
  const uint32_t call_pairs = (uint32_t)p.batch_size * 4u;
  auto load = [&](Batch *b, int par) {
    if (gpu_reader) {
      if (!LoadBatchGpu(ctx, &g1, se ? nullptr : &g2, sc ? &gb : nullptr, par, call_pairs, b, pairs, bc_len))
        Die(std::string("chromap-b200: the read files are not plain 4-line FASTQ (") + cmx_last_error(ctx) + "); rerun with --host-reader");
    } else LoadBatch(r1, r2, (uint32_t)p.batch_size, b, pairs || sam || paf, sc ? &rb : nullptr, bc_len, se, sam || paf);
  };
  open_all(gpu_reader);
  if (gpu_reader && !LoadBatchGpu(ctx, &g1, se ? nullptr : &g2, sc ? &gb : nullptr, parity, call_pairs, &cur, pairs, bc_len)) {
    fprintf(stderr, "Read files are not plain 4-line FASTQ (%s): using the host reader.\n", cmx_last_error(ctx));
    g1.Close(); g2.Close(); gb.Close();
    gpu_reader = false;
    open_all(false);
  }
  std::vector<cmx_pe_record> all, recs;
  bool recs_pinned = false;
  uint64_t n_pairs = 0, n_mapped = 0, n_unique = 0, n_cand = 0;
  const double t_map = Now();
  double t_calls = 0, t_collect = 0, t_wait = 0;
  int n_calls = 0;
  uint32_t read_id = 0;
  std::vector<std::string> all_names, all_names2;
  std::vector<cmx_sam_record> sam_recs, all_sam;
  std::string sam_s1, sam_q1, sam_s2, sam_q2;
  std::vector<uint64_t> sam_off1{0}, sam_off2{0};
  std::vector<uint16_t> paf_len1, paf_len2;
  std::vector<uint64_t> all_bc, bc_keys;
  uint64_t n_bc_in = 0, n_bc_cor = 0;
  if (!gpu_reader) load(&cur, parity);
  while (cur.n > 0) {
    cur.first_id = read_id;
    parity ^= 1;
    std::thread loader([&]() { load(&next, parity); });
    if (recs.size() < (size_t)cur.n * p.max_num_best_mappings) {  // the call's output buffer: page-locked, so the records come back at PCIe speed
      if (recs_pinned) { cmx_host_unregister(recs.data()); recs_pinned = false; }
      recs.resize((size_t)cur.n * p.max_num_best_mappings);
      recs_pinned = cmx_host_register(recs.data(), recs.size() * sizeof(cmx_pe_record)) == 0;
    }
    if (n_calls == 0 && gpu_reader && g1.file_bytes && g1.first_cut) {  // records of the whole run, from the first call's bytes per pair
      const double est = (double)g1.file_bytes / (double)g1.first_cut * cur.n * 1.02 + 1024;
      if (est < 4e9) all.reserve((size_t)est);
    }
    cmx_batch in{};
    if (cur.dev) in = cur.dev_in;
    else {
      in.n_pairs = cur.n; in.seq1 = cur.s1.data(); in.off1 = cur.o1.data(); in.seq2 = se ? nullptr : cur.s2.data(); in.off2 = se ? nullptr : cur.o2.data();
      if (sc) { in.bc_seq = cur.bc.data(); in.bc_qual = cur.bq.data(); in.bc_len = bc_len; }
    }
    in.first_read_id = cur.first_id;
    cmx_records out{};
    out.records = recs.data(); out.capacity = recs.size();
    if (sam) { sam_recs.resize(recs.size()); out.records = reinterpret_cast<cmx_pe_record *>(sam_recs.data()); }
    if (sc) { bc_keys.resize(recs.size()); out.barcode_keys = bc_keys.data(); }
    const double t0 = Now();
    if (cmx_map_batch_pe(ctx, &in, &out, nullptr)) Die(cmx_last_error(ctx));
    fprintf(stderr, se ? "Mapped %u reads in %.2fs.\n" : "Mapped %u read pairs in %.2fs.\n", cur.n, Now() - t0);
    t_calls += Now() - t0; ++n_calls;
    const double t_col0 = Now();
    if (sam) {
      all_sam.insert(all_sam.end(), sam_recs.begin(), sam_recs.begin() + out.n_records);
      for (uint32_t i = 0; i < cur.n; ++i) { sam_off1.push_back(sam_off1.back() + (cur.o1[i + 1] - cur.o1[i])); if (!se) sam_off2.push_back(sam_off2.back() + (cur.o2[i + 1] - cur.o2[i])); }
      sam_s1 += cur.s1; sam_q1 += cur.q1; sam_s2 += cur.s2; sam_q2 += cur.q2;
      all_names.insert(all_names.end(), cur.names1.begin(), cur.names1.end());
      all_names2.insert(all_names2.end(), cur.names2.begin(), cur.names2.end());
    } else all.insert(all.end(), recs.begin(), recs.begin() + out.n_records);
    if (paf) {
      for (uint32_t i = 0; i < cur.n; ++i) { paf_len1.push_back((uint16_t)(cur.o1[i + 1] - cur.o1[i])); if (!se) paf_len2.push_back((uint16_t)(cur.o2[i + 1] - cur.o2[i])); }
      all_names.insert(all_names.end(), cur.names1.begin(), cur.names1.end());
      all_names2.insert(all_names2.end(), cur.names2.begin(), cur.names2.end());
    }
    if (pairs) all_names.insert(all_names.end(), cur.names1.begin(), cur.names1.end());
    if (sc) { all_bc.insert(all_bc.end(), bc_keys.begin(), bc_keys.begin() + out.n_records); n_bc_in += out.n_barcodes_in_whitelist; n_bc_cor += out.n_barcodes_corrected; }
    n_pairs += cur.n; n_mapped += out.n_mapped_pairs; n_unique += out.n_uniquely_mapped_pairs; n_cand += out.n_candidates;
    read_id += cur.n;
    const double t_j0 = Now();
    t_collect += t_j0 - t_col0;
    loader.join();
    t_wait += Now() - t_j0;
    std::swap(cur, next);
  }
  fprintf(stderr, "Mapped all reads in %.2fs.\n", Now() - t_map);
  fprintf(stderr, "Mapping phase: %d calls: mapping %.3fs, collecting records %.3fs, waiting for the loader %.3fs (loader: reading + cutting %.3fs, parsing on the device %.3fs, first batch included).\n",
          n_calls, t_calls, t_collect, t_wait, g_t_fill, g_t_ingest);
  fprintf(stderr, "Number of reads: %llu.\nNumber of mapped reads: %llu.\nNumber of uniquely mapped reads: %llu.\nNumber of candidates: %llu.\n",
          (unsigned long long)((se ? 1 : 2) * n_pairs), (unsigned long long)((se ? 1 : 2) * n_mapped), (unsigned long long)((se ? 1 : 2) * n_unique),
          (unsigned long long)n_cand);
  uint64_t keep = 0;
  const double t_pp = Now();
  std::vector<const char *> names;
  for (const auto &s : ref.names) names.push_back(s.c_str());
  std::vector<char> text;
  int64_t bytes = 0;
  if (sam) {
    std::vector<const char *> n1, n2;
    for (const auto &x : all_names) n1.push_back(x.c_str());
    for (const auto &x : all_names2) n2.push_back(x.c_str());
    std::vector<uint32_t> lens;
    for (size_t i = 0; i + 1 < ref.offsets.size(); ++i) lens.push_back((uint32_t)(ref.offsets[i + 1] - ref.offsets[i]));
    cmx_read_set rs1{n1.data(), sam_s1.data(), sam_off1.data(), sam_q1.data()}, rs2{n2.data(), sam_s2.data(), sam_off2.data(), sam_q2.data()};
    bytes = cmx_format_sam(&p, names.data(), lens.data(), (uint32_t)names.size(), ref.concat.data(), ref.offsets.data(), all_sam.data(), all_sam.size(), &rs1,
                           se ? nullptr : &rs2, 0, nullptr, 0);
    if (bytes < 0) Die("chromap-b200: cmx_format_sam failed");
    text.resize((size_t)bytes + 1);
    cmx_format_sam(&p, names.data(), lens.data(), (uint32_t)names.size(), ref.concat.data(), ref.offsets.data(), all_sam.data(), all_sam.size(), &rs1,
                   se ? nullptr : &rs2, 0, text.data(), bytes);
    keep = 0;
    for (int64_t i = 0; i < bytes; ++i) keep += text[(size_t)i] == '\n';
    keep -= names.size();
  } else if (paf) {
    std::vector<const char *> n1, n2;
    for (const auto &x : all_names) n1.push_back(x.c_str());
    for (const auto &x : all_names2) n2.push_back(x.c_str());
    std::vector<uint32_t> lens;
    for (size_t i = 0; i + 1 < ref.offsets.size(); ++i) lens.push_back((uint32_t)(ref.offsets[i + 1] - ref.offsets[i]));
    bytes = cmx_format_paf(&p, names.data(), lens.data(), all.data(), all.size(), n1.data(), paf_len1.data(), se ? nullptr : n2.data(), se ? nullptr : paf_len2.data(), 0,
                           nullptr, 0);
    if (bytes < 0) Die("chromap-b200: cmx_format_paf failed");
    text.resize((size_t)bytes + 1);
    cmx_format_paf(&p, names.data(), lens.data(), all.data(), all.size(), n1.data(), paf_len1.data(), se ? nullptr : n2.data(), se ? nullptr : paf_len2.data(), 0,
                   text.data(), bytes);
    keep = 0;
    for (int64_t i = 0; i < bytes; ++i) keep += text[(size_t)i] == '\n';
    if (!se) keep /= 2;
  } else if (pairs) {
    cmx_pairs_record *pr = reinterpret_cast<cmx_pairs_record *>(all.data());
    if (cmx_postprocess_gpu(ctx, pr, nullptr, all.size(), &keep) && cmx_postprocess_pairs(ctx, pr, all.size(), &keep)) Die(cmx_last_error(ctx));
    std::vector<const char *> rn;
    for (const auto &s : all_names) rn.push_back(s.c_str());
    std::vector<uint32_t> lens;
    for (size_t i = 0; i + 1 < ref.offsets.size(); ++i) lens.push_back((uint32_t)(ref.offsets[i + 1] - ref.offsets[i]));
    bytes = cmx_format_pairs_gpu(ctx, names.data(), lens.data(), (uint32_t)names.size(), pr, keep, rn.data(), rn.size(), 0, nullptr, 0);
    if (bytes >= 0) { text.resize((size_t)bytes + 1); bytes = cmx_format_pairs_gpu(ctx, names.data(), lens.data(), (uint32_t)names.size(), pr, keep, rn.data(), rn.size(), 0, text.data(), bytes); }
    if (bytes < 0) {
      bytes = cmx_format_pairs(names.data(), lens.data(), (uint32_t)names.size(), pr, keep, rn.data(), 0, nullptr, 0);
      text.resize((size_t)bytes + 1);
      cmx_format_pairs(names.data(), lens.data(), (uint32_t)names.size(), pr, keep, rn.data(), 0, text.data(), bytes);
    }
  } else if (sc) {
    if (cmx_postprocess_gpu(ctx, all.data(), all_bc.data(), all.size(), &keep) && cmx_postprocess_bc(ctx, all.data(), all_bc.data(), all.size(), &keep)) Die(cmx_last_error(ctx));
    bytes = cmx_format_bed_gpu(ctx, names.data(), all.data(), all_bc.data(), keep, bc_len, nullptr, 0);  // text written on the device
    if (bytes >= 0) { text.resize((size_t)bytes + 1); bytes = cmx_format_bed_gpu(ctx, names.data(), all.data(), all_bc.data(), keep, bc_len, text.data(), bytes); }
    if (bytes < 0) {
      bytes = cmx_format_bed_bc(names.data(), all.data(), all_bc.data(), keep, bc_len, nullptr, 0);
      text.resize((size_t)bytes + 1);
      cmx_format_bed_bc(names.data(), all.data(), all_bc.data(), keep, bc_len, text.data(), bytes);
    }
    fprintf(stderr, "Number of barcodes in whitelist: %llu.\nNumber of corrected barcodes: %llu.\n", (unsigned long long)n_bc_in, (unsigned long long)n_bc_cor);
  } else {
    // sort / dedup / filter on the device; the host routine only if the records do not fit beside the index
    if (cmx_postprocess_gpu(ctx, all.data(), nullptr, all.size(), &keep) && cmx_postprocess(ctx, all.data(), all.size(), &keep)) Die(cmx_last_error(ctx));
    if (tagalign && !se) {  // PairedTagAlign: two lines per pair (host formatter); single-end TagAlign is the BED text
      bytes = cmx_format_tagalign(names.data(), all.data(), keep, nullptr, 0);
      text.resize((size_t)bytes + 1);
      cmx_format_tagalign(names.data(), all.data(), keep, text.data(), bytes);
    } else
    bytes = cmx_format_bed_gpu(ctx, names.data(), all.data(), nullptr, keep, 0, nullptr, 0);  // text written on the device
    if (!(tagalign && !se) && bytes >= 0) { text.resize((size_t)bytes + 1); bytes = cmx_format_bed_gpu(ctx, names.data(), all.data(), nullptr, keep, 0, text.data(), bytes); }
    if (bytes < 0) {
      bytes = cmx_format_bed(names.data(), all.data(), keep, nullptr, 0);
      text.resize((size_t)bytes + 1);
      cmx_format_bed(names.data(), all.data(), keep, text.data(), bytes);
    }
  }
  FILE *fo = fopen(out_path.c_str(), "wb");
  if (!fo) Die("Cannot open output file " + out_path);
  fwrite(text.data(), 1, (size_t)bytes, fo);
  fclose(fo);
  fprintf(stderr, "Sorted, deduped and outputed mappings in %.2fs.\nNumber of output mappings (passed filters): %llu\nTotal time: %.2fs.\n", Now() - t_pp,
          (unsigned long long)keep, Now() - t_start);
  cmx_destroy(ctx);
  return 0;
}
