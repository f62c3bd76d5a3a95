#include "seqio.h"

#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstring>

#include <cctype>
#include <cstdio>

namespace cmxhost {

bool SeqReader::Open(const std::string &path) {
  Close();
  f_ = gzopen(path.c_str(), "r");
  if (!f_) return false;
  gzbuffer(f_, 1 << 20);
  buf_.resize(1 << 20);
  pos_ = end_ = 0;
  eof_ = false;
  pending_ = 0;
  return true;
}

void SeqReader::Close() {
  if (f_) gzclose(f_);
  f_ = nullptr;
}

int SeqReader::GetC() {
  if (pos_ >= end_) {
    if (eof_) return -1;
    const int n = gzread(f_, buf_.data(), (unsigned)buf_.size());
    if (n <= 0) { eof_ = true; return -1; }
    pos_ = 0;
    end_ = (size_t)n;
  }
  return buf_[pos_++];
}

bool SeqReader::GetLine(std::string *s) {
  s->clear();
  bool any = false;
  for (;;) {
    if (pos_ >= end_) {
      if (eof_) break;
      const int n = gzread(f_, buf_.data(), (unsigned)buf_.size());
      if (n <= 0) { eof_ = true; break; }
      pos_ = 0;
      end_ = (size_t)n;
    }
    any = true;
    const void *nlp = memchr(buf_.data() + pos_, '\n', end_ - pos_);
    const size_t i = nlp ? (size_t)((const unsigned char *)nlp - buf_.data()) : end_;
    s->append((const char *)buf_.data() + pos_, i - pos_);
    if (i < end_) { pos_ = i + 1; break; }
    pos_ = end_;
  }
  if (s->size() > 1 && s->back() == '\r') s->pop_back();  // kseq.h:141 strips '\r' only when the line has more
  return any;
}

bool SeqReader::Next(std::string *name, std::string *seq, std::string *qual) {
  name->clear(); seq->clear(); qual->clear();
  int c;
  if (pending_ == 0) {
    while ((c = GetC()) != -1 && c != '>' && c != '@') {}
    if (c == -1) return false;
  }
  pending_ = 0;
  std::string line;
  GetLine(&line);
  size_t sp = 0;
  while (sp < line.size() && !isspace((unsigned char)line[sp])) ++sp;
  name->assign(line, 0, sp);
  while ((c = GetC()) != -1 && c != '>' && c != '+' && c != '@') {
    if (c == '\n') continue;
    seq->push_back((char)c);
    GetLine(&line);
    seq->append(line);
    if (seq->size() > 1 && seq->back() == '\r') seq->pop_back();  // kseq.h:141 tests the whole accumulated string (append mode)
  }
  if (c == '>' || c == '@') pending_ = c;
  if (c != '+') return true;
  GetLine(&line);  // rest of the '+' line
  while (qual->size() < seq->size()) {
    if (!GetLine(&line)) break;
    qual->append(line);
    if (qual->size() > 1 && qual->back() == '\r') qual->pop_back();
  }
  // kseq.h:213-218: a missing or differently long quality string is an error (-2); the reference's loader then stops with
  // "Didn't reach the end of sequence file, which might be corrupted!" (sequence_batch.cc:46-55)
  if (qual->size() != seq->size()) corrupted_ = true;
  return !corrupted_;
}

// One line appended to *dst (no intermediate string); the same '\r' rule as GetLine.  Returns false at end of file.
bool SeqReader::AppendLine(std::string *dst) {
  const size_t before = dst->size();
  bool any = false;
  for (;;) {
    if (pos_ >= end_) {
      if (eof_) break;
      const int n = gzread(f_, buf_.data(), (unsigned)buf_.size());
      if (n <= 0) { eof_ = true; break; }
      pos_ = 0;
      end_ = (size_t)n;
    }
    any = true;
    const void *nlp = memchr(buf_.data() + pos_, '\n', end_ - pos_);
    const size_t i = nlp ? (size_t)((const unsigned char *)nlp - buf_.data()) : end_;
    dst->append((const char *)buf_.data() + pos_, i - pos_);
    if (i < end_) { pos_ = i + 1; break; }
    pos_ = end_;
  }
  if (dst->size() - before > 1 && dst->back() == '\r') dst->pop_back();
  return any;
}

// Next() for a reference: the sequence goes straight to the end of *dst (3 Gbp pass through one copy instead of three).
bool SeqReader::NextAppend(std::string *name, std::string *dst) {
  name->clear();
  int c;
  if (pending_ == 0) {
    while ((c = GetC()) != -1 && c != '>' && c != '@') {}
    if (c == -1) return false;
  }
  pending_ = 0;
  std::string line;
  GetLine(&line);
  size_t sp = 0;
  while (sp < line.size() && !isspace((unsigned char)line[sp])) ++sp;
  name->assign(line, 0, sp);
  const size_t start = dst->size();
  while ((c = GetC()) != -1 && c != '>' && c != '+' && c != '@') {
    if (c == '\n') continue;
    dst->push_back((char)c);
    AppendLine(dst);
    if (dst->size() - start > 1 && dst->back() == '\r') dst->pop_back();
  }
  if (c == '>' || c == '@') pending_ = c;
  if (c != '+') return true;
  GetLine(&line);  // rest of the '+' line; the quality is consumed by length (kseq.h:205-218)
  const size_t len = dst->size() - start;
  size_t q = 0;
  while (q < len) {
    if (!GetLine(&line)) break;
    q += line.size();
    if (q > 1 && !line.empty() && line.back() == '\r') --q;
  }
  if (q != len) corrupted_ = true;
  return !corrupted_;
}

bool Reference::Load(const std::string &path) {
  SeqReader rd;
  if (!rd.Open(path)) return false;
  names.clear(); concat.clear(); offsets.assign(1, 0);
  struct stat st;
  if (stat(path.c_str(), &st) == 0 && st.st_size > 0) concat.reserve((size_t)st.st_size);  // a plain file: never grown again
  std::string n;
  while (rd.NextAppend(&n, &concat)) {
    if (concat.size() == offsets.back()) continue;  // empty records are skipped (sequence_batch.cc:84-118)
    names.push_back(n);
    offsets.push_back(concat.size());
  }
  concat.resize(offsets.back());  // (a record cut short by a corrupted file leaves nothing behind)
  return !names.empty();
}

bool IndexFile::Load(const std::string &path) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) return false;
  uint32_t lookup_size = 0, n_occ = 0;
  bool ok = fread(&k, 4, 1, f) == 1 && fread(&w, 4, 1, f) == 1 && fread(&lookup_size, 4, 1, f) == 1 &&
            fread(&n_buckets, 4, 1, f) == 1 && fread(&size, 4, 1, f) == 1 && fread(&n_occupied, 4, 1, f) == 1 &&
            fread(&upper_bound, 4, 1, f) == 1;
  if (ok && n_buckets) {
    const size_t nf = n_buckets < 16 ? 1 : n_buckets >> 4;
    flags.resize(nf); keys.resize(n_buckets); vals.resize(n_buckets);
    ok = fread(flags.data(), 4, nf, f) == nf && fread(keys.data(), 8, n_buckets, f) == n_buckets &&
         fread(vals.data(), 8, n_buckets, f) == n_buckets;
  }
  ok = ok && fread(&n_occ, 4, 1, f) == 1;
  if (ok && n_occ) { occ.resize(n_occ); ok = fread(occ.data(), 8, n_occ, f) == n_occ; }
  fclose(f);
  return ok;
}

bool IndexMap::Open(const std::string &path) {
  Close();
  const int fd = open(path.c_str(), O_RDONLY);
  if (fd < 0) return false;
  struct stat st;
  if (fstat(fd, &st) != 0 || st.st_size < 32) { close(fd); return false; }
  bytes_ = (size_t)st.st_size;
  base_ = mmap(nullptr, bytes_, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (base_ == MAP_FAILED) { base_ = nullptr; return false; }
  madvise(base_, bytes_, MADV_SEQUENTIAL);
  madvise(base_, bytes_, MADV_WILLNEED);
  const unsigned char *p = (const unsigned char *)base_;
  uint32_t h[7];
  memcpy(h, p, 28);
  k = (int)h[0]; w = (int)h[1]; n_buckets = h[3]; size = h[4];
  const size_t nf = n_buckets < 16 ? 1 : n_buckets >> 4;
  size_t off = 28;
  const size_t need = off + (n_buckets ? nf * 4 + (size_t)n_buckets * 16 : 0) + 4;
  if (need > bytes_) { Close(); return false; }
  if (n_buckets) {
    flags = (const uint32_t *)(p + off); off += nf * 4;
    keys = (const uint64_t *)(p + off); off += (size_t)n_buckets * 8;
    vals = (const uint64_t *)(p + off); off += (size_t)n_buckets * 8;
  }
  memcpy(&n_occ, p + off, 4); off += 4;
  if (off + (size_t)n_occ * 8 > bytes_) { Close(); return false; }
  occ = (const uint64_t *)(p + off);
  return true;
}
void IndexMap::Close() {
  if (base_) munmap(base_, bytes_);
  base_ = nullptr; bytes_ = 0; flags = nullptr; keys = vals = occ = nullptr;
}

bool IndexFile::Save(const std::string &path) const {
  FILE *f = fopen(path.c_str(), "wb");
  if (!f) return false;
  const uint32_t n_occ = (uint32_t)occ.size();
  bool ok = fwrite(&k, 4, 1, f) == 1 && fwrite(&w, 4, 1, f) == 1 && fwrite(&size, 4, 1, f) == 1 &&
            fwrite(&n_buckets, 4, 1, f) == 1 && fwrite(&size, 4, 1, f) == 1 && fwrite(&n_occupied, 4, 1, f) == 1 &&
            fwrite(&upper_bound, 4, 1, f) == 1;
  if (ok && n_buckets)
    ok = fwrite(flags.data(), 4, flags.size(), f) == flags.size() && fwrite(keys.data(), 8, n_buckets, f) == n_buckets &&
         fwrite(vals.data(), 8, n_buckets, f) == n_buckets;
  ok = ok && fwrite(&n_occ, 4, 1, f) == 1;
  if (ok && n_occ) ok = fwrite(occ.data(), 8, n_occ, f) == n_occ;
  fclose(f);
  return ok;
}

}  // namespace cmxhost
