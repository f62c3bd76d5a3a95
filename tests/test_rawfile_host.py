"""The CLI's raw FASTQ reader (chromap_b200/csrc/host/cli.cc, RawFile: threaded pread + bulk newline counting + cut at whole
4-line records) compiled for the HOST against a line-by-line walk over the same bytes: every Fill / Consume sequence must
hand out the same byte ranges and record counts, whatever the chunk size, line lengths, a missing final newline or a
trailing partial record.  No device needed (the page-locking call fails softly without one)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#define main cli_main
#include "%(root)s/chromap_b200/csrc/host/cli.cc"
#undef main
#include <random>
static std::string make_text(std::mt19937 &g, int n_rec, int tail_lines, bool final_newline) {
  std::string t;
  auto line = [&](int lo, int hi) { const int L = lo + (int)(g() %% (unsigned)(hi - lo + 1)); for (int i = 0; i < L; ++i) t.push_back("ACGTN"[g() %% 5]); t.push_back('\n'); };
  for (int r = 0; r < n_rec; ++r) { t.push_back('@'); line(0, 20); line(0, 200); t += "+\n"; line(0, 200); }
  for (int q = 0; q < tail_lines; ++q) line(0, 30);
  if (!final_newline && !t.empty() && t.back() == '\n') t.pop_back();
  return t;
}
int main(int argc, char **argv) {
  std::mt19937 g(5);
  long bad = 0, cases = 0;
  for (int it = 0; it < 300; ++it) {
    const int n_rec = (int)(g() %% 400), tail = (int)(g() %% 4);
    const bool fin = g() %% 2;
    std::string text = make_text(g, n_rec, tail, fin);
    const std::string path = argc > 1 ? argv[1] : "/tmp/cmx_rawfile_test.fq";
    FILE *f = fopen(path.c_str(), "wb"); fwrite(text.data(), 1, text.size(), f); fclose(f);
    std::string ref = text;                       // what the reader works on: the text with its last newline made up
    if (!ref.empty() && ref.back() != '\n') ref.push_back('\n');
    RawFile rf;
    if (!rf.Open(path)) { printf("open failed\n"); return 2; }
    rf.chunk = it %% 3 == 0 ? 0 : 64 + g() %% 8192;   // 0: the production chunk size (clamped to what is left of the file)
    if (it %% 6 == 0) rf.Prepare(1 + g() %% 500);       // the start-up sizing from the head of the file
    size_t pos = 0;                               // bytes handed out so far
    for (int call = 0; call < 1000; ++call) {
      const uint32_t want = 1 + g() %% 97;
      uint32_t n = 0;
      const uint64_t bytes = rf.Fill(want, &n);
      if (call == 0 && it %% 2 == 0) {  // the start-up thread fills the first call ahead of time: asking again must give the same answer
        uint32_t n2 = 0;
        const uint64_t b2 = rf.Fill(want, &n2);
        if (b2 != bytes || n2 != n) { if (bad < 5) printf("REFILL it=%%d %%llu/%%llu %%u/%%u\n", it, (unsigned long long)b2, (unsigned long long)bytes, n2, n); ++bad; }
      }
      // the walk: up to `want` records = 4 lines each starting at pos
      size_t p = pos; uint32_t rn = 0;
      while (rn < want) {
        size_t q = p; int l = 0;
        while (l < 4) { const size_t e = ref.find('\n', q); if (e == std::string::npos) break; q = e + 1; ++l; }
        if (l < 4) break;
        p = q; ++rn;
      }
      ++cases;
      if (n != rn || bytes != p - pos || memcmp(rf.buf.data(), ref.data() + pos, bytes)) {
        if (bad < 5) printf("MISMATCH it=%%d call=%%d want=%%u got n=%%u bytes=%%llu expect n=%%u bytes=%%zu\n", it, call, want, n, (unsigned long long)bytes, rn, p - pos);
        ++bad; break;
      }
      rf.Consume(bytes);
      pos = p;
      if (n == 0) break;
    }
    // what stays behind is exactly the unfinished record (if any)
    if (rf.have != ref.size() - pos) { if (bad < 5) printf("LEFTOVER it=%%d have=%%zu expect=%%zu\n", it, rf.have, ref.size() - pos); ++bad; }
  }
  printf("fill_calls=%%ld bad=%%ld\n", cases, bad);
  return bad != 0;
}
'''


def test_raw_fastq_reader_cuts_like_a_line_walk(tmp_path):
    src = tmp_path / "t.cc"
    src.write_text(SRC % dict(root=ROOT))
    exe = tmp_path / "t"
    lib = os.path.join(ROOT, "chromap_b200")
    subprocess.run(["g++", "-std=c++17", "-O2", "-o", str(exe), str(src), os.path.join(ROOT, "chromap_b200/csrc/host/seqio.cc"), "-L" + lib, "-lchromap_b200",
                    "-Wl,-rpath," + lib, "-lz", "-lpthread"], check=True)
    r = subprocess.run([str(exe), str(tmp_path / "reads.fq")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "bad=0" in r.stdout


REF_SRC = r'''
#include "%(root)s/chromap_b200/csrc/host/seqio.h"
#include <cstdio>
#include <cstring>
#include <random>
using namespace cmxhost;
int main(int argc, char **argv) {
  std::mt19937 g(9);
  long bad = 0;
  for (int it = 0; it < 400; ++it) {
    std::string t;
    const int n_rec = 1 + (int)(g() %% 12);
    const bool crlf = g() %% 4 == 0, fastq = g() %% 6 == 0;
    auto nl = [&]() { if (crlf) t.push_back('\r'); t.push_back('\n'); };
    if (g() %% 3 == 0) { t += "junk before the first record"; nl(); }
    for (int r = 0; r < n_rec; ++r) {
      t.push_back(fastq ? '@' : '>'); t += "seq" + std::to_string(r); if (g() %% 2) t += " a comment\twith tabs"; nl();
      const int n_lines = (int)(g() %% 5);        // 0 lines: an empty record (skipped)
      std::string seq;
      for (int l = 0; l < n_lines; ++l) {
        const int L = (int)(g() %% 70);
        for (int i = 0; i < L; ++i) { const char c = "ACGTNacgtn"[g() %% 10]; t.push_back(c); seq.push_back(c); }
        if (L == 1 && crlf && g() %% 2) { }      // (a one-base line followed by CR LF: the CR rule's edge)
        nl();
        if (g() %% 9 == 0) nl();                  // blank line inside a record
      }
      if (fastq) { t += "+"; nl(); for (size_t i = 0; i < seq.size(); ++i) t.push_back('I'); nl(); }
    }
    if (g() %% 5 == 0 && !t.empty() && t.back() == '\n') t.pop_back();
    const std::string path = argc > 1 ? argv[1] : "/tmp/cmx_ref_test.fa";
    FILE *f = fopen(path.c_str(), "wb"); fwrite(t.data(), 1, t.size(), f); fclose(f);
    // the record-at-a-time reader (the one the read files go through) is the specification
    std::vector<std::string> names; std::string concat; std::vector<uint64_t> offs{0};
    { SeqReader rd; rd.Open(path); std::string n, s, q; while (rd.Next(&n, &s, &q)) { if (s.empty()) continue; names.push_back(n); concat += s; offs.push_back(concat.size()); } }
    Reference ref;
    const bool ok = ref.Load(path);
    if (ok != !names.empty() || (ok && (ref.names != names || ref.concat != concat || ref.offsets != offs))) {
      if (bad < 5) printf("MISMATCH it=%%d records %%zu/%%zu bytes %%zu/%%zu\n", it, ref.names.size(), names.size(), ref.concat.size(), concat.size());
      ++bad;
    }
  }
  printf("bad=%%ld\n", bad);
  return bad != 0;
}
'''


def test_reference_loader_equals_record_reader(tmp_path):
    src = tmp_path / "r.cc"
    src.write_text(REF_SRC % dict(root=ROOT))
    exe = tmp_path / "r"
    subprocess.run(["g++", "-std=c++17", "-O2", "-o", str(exe), str(src), os.path.join(ROOT, "chromap_b200/csrc/host/seqio.cc"), "-lz"], check=True)
    r = subprocess.run([str(exe), str(tmp_path / "ref.fa")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "bad=0" in r.stdout
