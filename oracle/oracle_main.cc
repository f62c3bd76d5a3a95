// TEST INFRASTRUCTURE ONLY — command-line front end of the CPU oracle (see oracle_chromap.h).
// Usage mirrors the reference CLI subset: oracle_map [--preset P] [-e N] [-q N] [-l N] [-t N]
//   [--trim-adapters] [--remove-pcr-duplicates] [--Tn5-shift] [--low-mem] -x idx -r ref -1 r1 -2 r2 -o out
//   or: oracle_map -i -r ref.fa -o out.index [-k 17 -w 7]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "oracle_chromap.h"

int main(int argc, char **argv) {
  orc_params p;
  orc_default_params(&p);
  std::string idx, ref, r1, r2, out, preset, bc, wl;
  int threads = 1, k = 17, w = 7;
  bool build = false;
  for (int i = 1; i < argc; ++i)
    if (!strcmp(argv[i], "--preset") && i + 1 < argc) preset = argv[++i];
  if (orc_apply_preset(&p, preset.c_str()) != 0) { fprintf(stderr, "Unrecognized preset parameters %s\n", preset.c_str()); return 255; }
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    auto next = [&]() -> const char * { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(255); } return argv[++i]; };
    if (a == "--preset") next();
    else if (a == "-i") build = true;
    else if (a == "-x") idx = next();
    else if (a == "-r") ref = next();
    else if (a == "-1") r1 = next();
    else if (a == "-2") r2 = next();
    else if (a == "-o") out = next();
    else if (a == "-b") bc = next();
    else if (a == "--barcode-whitelist") wl = next();
    else if (a == "-t") threads = atoi(next());
    else if (a == "-k") k = atoi(next());
    else if (a == "-w") w = atoi(next());
    else if (a == "-e") p.error_threshold = atoi(next());
    else if (a == "-q") p.mapq_threshold = atoi(next());
    else if (a == "-l") p.max_insert_size = atoi(next());
    else if (a == "--trim-adapters") p.trim_adapters = 1;
    else if (a == "--remove-pcr-duplicates") p.remove_pcr_duplicates = 1;
    else if (a == "--Tn5-shift") p.tn5_shift = 1;
    else if (a == "--low-mem") p.low_memory_mode = 1;
    else if (a == "--split-alignment") p.split_alignment = 1;
    else if (a == "--pairs") p.output_format = 5;
    else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 255; }
  }
  if (build) {
    orc_reference *r = orc_reference_load(ref.c_str());
    if (!r) { fprintf(stderr, "Cannot find sequence file %s\n", ref.c_str()); return 255; }
    orc_index *ix = orc_index_build(r, k, w);
    return orc_index_save(ix, out.c_str()) == 0 ? 0 : 255;
  }
  if (!bc.empty()) {
    uint64_t st[2] = {0, 0};
    int rc = orc_run_files_bc(&p, idx.c_str(), ref.c_str(), r1.c_str(), r2.c_str(), bc.c_str(), wl.c_str(), out.c_str(), threads, st);
    if (rc != 0) { fprintf(stderr, "oracle_map failed (%d)\n", rc); return 255; }
    fprintf(stderr, "Number of barcodes in whitelist: %llu.\nNumber of corrected barcodes: %llu.\n", (unsigned long long)st[0], (unsigned long long)st[1]);
    return 0;
  }
  double secs = 0;
  uint64_t n = 0;
  int rc = orc_run_files(&p, idx.c_str(), ref.c_str(), r1.c_str(), r2.c_str(), out.c_str(), threads, &secs, &n);
  if (rc != 0) { fprintf(stderr, "oracle_map failed (%d)\n", rc); return 255; }
  fprintf(stderr, "Mapped all reads in %.2fs.\n", secs);
  fprintf(stderr, "Number of pairs: %llu.\n", (unsigned long long)n);
  return 0;
}
