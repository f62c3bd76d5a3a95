// chromap_b200 host side — FASTA/FASTQ reader with the behaviour of the reference's kseq.h
// (kseq.h:177-222): record starts at '>' or '@'; name = header up to the first whitespace; sequence = all
// following lines (newline and a trailing '\r' stripped, every other byte kept) until a line starting
// with '>', '+' or '@'; FASTQ quality is consumed by length.  Plain or gzip (zlib), like gzopen.
#pragma once
#include <stdint.h>
#include <zlib.h>

#include <string>
#include <vector>

namespace cmxhost {

class SeqReader {
 public:
  bool Open(const std::string &path);
  void Close();
  // false at end of file or on a corrupted record (Corrupted() tells which).  `qual` stays empty for FASTA.
  bool Next(std::string *name, std::string *seq, std::string *qual);
  // the same record with its sequence appended to *dst instead of returned (reference loading)
  bool NextAppend(std::string *name, std::string *dst);
  bool Corrupted() const { return corrupted_; }
  ~SeqReader() { Close(); }

 private:
  int GetC();
  bool GetLine(std::string *s);
  bool AppendLine(std::string *dst);
  gzFile f_ = nullptr;
  std::vector<unsigned char> buf_;
  size_t pos_ = 0, end_ = 0;
  bool eof_ = false;
  int pending_ = 0;  // header character already consumed
  bool corrupted_ = false;
};

struct Reference {
  std::vector<std::string> names;
  std::string concat;             // all sequences back to back, bytes as loaded
  std::vector<uint64_t> offsets;  // n + 1
  bool Load(const std::string &path);  // SequenceBatch::LoadAllSequences (sequence_batch.cc:84-118): empty records are skipped
};

// Index file, reference layout (index.cc:91-169, khash.h:358-386).
struct IndexFile {
  int k = 0, w = 0;
  uint32_t n_buckets = 0, size = 0, n_occupied = 0, upper_bound = 0;
  std::vector<uint32_t> flags;
  std::vector<uint64_t> keys, vals, occ;
  bool Load(const std::string &path);
  bool Save(const std::string &path) const;
};

// The same file mapped read-only instead of copied: the arrays are handed to cmx_upload_index where they lie in the page
// cache (a 3 Gbp index is 17 GB of hash-table arrays; reading them into vectors first doubles the memory traffic of start-up).
struct IndexMap {
  int k = 0, w = 0;
  uint32_t n_buckets = 0, size = 0;
  const uint32_t *flags = nullptr;
  const uint64_t *keys = nullptr, *vals = nullptr, *occ = nullptr;  // 4-byte aligned only: to be copied, not dereferenced
  uint32_t n_occ = 0;
  bool Open(const std::string &path);
  void Close();
  ~IndexMap() { Close(); }

 private:
  void *base_ = nullptr;
  size_t bytes_ = 0;
};

}  // namespace cmxhost
