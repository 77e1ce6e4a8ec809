// Host-side byte-level BPE encoder of the CLIP text transform (text half): the merge loop of
// torchmultimodal/transforms/clip_transform.py:82-190 (CLIPBPETokenizer) as native code, called once per batch with the
// regex-split, lower-cased word pieces.  No device work: tokenisation is string processing; what the GPU needs is the
// [B, 77] id tensor, which the Python wrapper builds and uploads.  (Compiled by nvcc as plain host C++.)
//
// Semantics follow the reference's string model so that ids agree even on odd merge files:
//   alphabet  : every byte maps to one printable code point (bytes_to_unicode, :31-55); symbols are strings over it, the
//               last symbol of a word carries the suffix "</w>";
//   vocabulary: 256 byte symbols, the same 256 + "</w>", one entry per merge (left + right), then bos, eos — a later
//               duplicate string takes the later index (dict comprehension, :129);
//   merging   : repeatedly take the adjacent pair of lowest rank and merge its occurrences left to right (:140-172);
//   words equal to the bos / eos token strings encode to those ids (the reference seeds its cache with them, :131).
// Own data structures: ranks in a hash map keyed by "left\x01right", a per-encoder word cache, symbols merged in place.
#include <cstdint>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "mmb200_internal.h"

namespace {

void append_utf8(std::string& s, uint32_t cp) {
  if (cp < 0x80) {
    s.push_back((char)cp);
  } else if (cp < 0x800) {
    s.push_back((char)(0xC0 | (cp >> 6)));
    s.push_back((char)(0x80 | (cp & 0x3F)));
  } else {
    s.push_back((char)(0xE0 | (cp >> 12)));
    s.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
    s.push_back((char)(0x80 | (cp & 0x3F)));
  }
}

struct Bpe {
  std::string alphabet[256];                       // byte -> its alphabet string (UTF-8)
  std::unordered_map<std::string, int> rank;       // "left\x01right" -> merge rank
  std::unordered_map<std::string, int> vocab;      // symbol string -> id
  std::unordered_map<std::string, std::vector<int>> cache;   // raw word bytes -> ids
  std::string bos, eos;
  int n_vocab = 0;
  std::mutex mu;
};

void build_alphabet(Bpe& b, std::vector<std::string>& order) {
  // printable bytes keep their code point; the others get 256, 257, ... in increasing byte order
  bool printable[256] = {};
  for (int c = '!'; c <= '~'; ++c) printable[c] = true;
  for (int c = 0xA1; c <= 0xAC; ++c) printable[c] = true;
  for (int c = 0xAE; c <= 0xFF; ++c) printable[c] = true;
  std::vector<int> bytes_in_order;
  for (int c = 0; c < 256; ++c)
    if (printable[c]) bytes_in_order.push_back(c);
  uint32_t next = 256;
  std::vector<int> rest;
  for (int c = 0; c < 256; ++c)
    if (!printable[c]) rest.push_back(c);
  for (int c : bytes_in_order) { b.alphabet[c].clear(); append_utf8(b.alphabet[c], (uint32_t)c); }
  for (int c : rest) { b.alphabet[c].clear(); append_utf8(b.alphabet[c], next++); }
  for (int c : bytes_in_order) order.push_back(b.alphabet[c]);
  for (int c : rest) order.push_back(b.alphabet[c]);
}

std::vector<std::string> split_ws(const std::string& line) {
  std::vector<std::string> out;
  size_t i = 0;
  while (i < line.size()) {
    while (i < line.size() && (line[i] == ' ' || line[i] == '\t' || line[i] == '\r')) ++i;
    size_t j = i;
    while (j < line.size() && line[j] != ' ' && line[j] != '\t' && line[j] != '\r') ++j;
    if (j > i) out.emplace_back(line.substr(i, j - i));
    i = j;
  }
  return out;
}

void encode_word(Bpe& b, const char* w, size_t n, std::vector<int>& out) {
  std::string key(w, n);
  auto hit = b.cache.find(key);
  if (hit != b.cache.end()) { out = hit->second; return; }
  std::vector<std::string> sym;
  sym.reserve(n);
  std::string whole;
  for (size_t i = 0; i < n; ++i) { sym.push_back(b.alphabet[(unsigned char)w[i]]); whole += sym.back(); }
  out.clear();
  if (whole == b.bos || whole == b.eos) {
    out.push_back(b.vocab.at(whole));
  } else if (!sym.empty()) {
    sym.back() += "</w>";
    while (sym.size() > 1) {
      int best = INT32_MAX;
      size_t at = 0;
      for (size_t i = 0; i + 1 < sym.size(); ++i) {
        auto r = b.rank.find(sym[i] + '\x01' + sym[i + 1]);
        if (r != b.rank.end() && r->second < best) { best = r->second; at = i; }
      }
      if (best == INT32_MAX) break;
      const std::string left = sym[at], right = sym[at + 1];
      size_t wpos = 0;
      for (size_t i = 0; i < sym.size();) {   // merge every occurrence of (left, right), left to right, in place
        if (i + 1 < sym.size() && sym[i] == left && sym[i + 1] == right) { sym[wpos++] = left + right; i += 2; }
        else { if (wpos != i) sym[wpos] = std::move(sym[i]); ++wpos; ++i; }
      }
      sym.resize(wpos);
    }
    for (const std::string& s : sym) out.push_back(b.vocab.at(s));   // every byte symbol is in the vocabulary
  }
  b.cache.emplace(std::move(key), out);
}

}  // namespace

// merges_utf8: the whole merges file (first line = header, dropped, as the reference does with `split("\n")[1:]`);
// num_merges <= 0: all lines that follow.  bos / eos: the special-token strings.
extern "C" int mmb_bpe_create(const char* merges_utf8, long long n_bytes, int num_merges, const char* bos, const char* eos,
                              void** handle, int* vocab_size) {
  if (!merges_utf8 || n_bytes < 0 || !bos || !eos || !handle) return MMB_ERR_ARG;
  Bpe* b = new Bpe();
  std::vector<std::string> vocab;
  build_alphabet(*b, vocab);
  for (int i = 0; i < 256; ++i) vocab.push_back(vocab[i] + "</w>");
  const std::string text(merges_utf8, (size_t)n_bytes);
  std::vector<std::string> lines;
  size_t pos = 0;
  while (true) {   // Python str.split("\n"): n separators -> n + 1 pieces (a trailing newline yields a last empty piece)
    size_t nl = text.find('\n', pos);
    if (nl == std::string::npos) { lines.emplace_back(text.substr(pos)); break; }
    lines.emplace_back(text.substr(pos, nl - pos));
    pos = nl + 1;
  }
  const size_t avail = lines.size() > 0 ? lines.size() - 1 : 0;
  size_t take = (num_merges <= 0) ? avail : ((size_t)num_merges < avail ? (size_t)num_merges : avail);
  for (size_t k = 0; k < take; ++k) {
    const std::vector<std::string> parts = split_ws(lines[1 + k]);
    std::string joined;
    for (const std::string& p : parts) joined += p;
    if (parts.size() == 2) b->rank[parts[0] + '\x01' + parts[1]] = (int)k;   // a later duplicate pair keeps the later rank
    vocab.push_back(joined);                                                  // every line takes a vocabulary slot
  }
  vocab.emplace_back(bos);
  vocab.emplace_back(eos);
  for (size_t i = 0; i < vocab.size(); ++i) b->vocab[vocab[i]] = (int)i;      // later duplicates win, as in the reference
  b->n_vocab = (int)vocab.size();
  b->bos = bos;
  b->eos = eos;
  *handle = b;
  if (vocab_size) *vocab_size = b->n_vocab;
  return MMB_OK;
}

// words: concatenated UTF-8 bytes of n_words word pieces, piece i = [offsets[i], offsets[i+1]).  out_ids: capacity
// offsets[n_words] (a piece never yields more ids than it has bytes); out_counts[i] = ids of piece i.
extern "C" int mmb_bpe_encode(void* handle, const char* words, const long long* offsets, int n_words, int* out_ids,
                              int* out_counts) {
  if (!handle || n_words < 0 || (n_words > 0 && (!words || !offsets || !out_ids || !out_counts))) return MMB_ERR_ARG;
  Bpe* b = static_cast<Bpe*>(handle);
  std::lock_guard<std::mutex> lock(b->mu);
  std::vector<int> ids;
  long long w = 0;
  try {
    for (int i = 0; i < n_words; ++i) {
      const long long lo = offsets[i], hi = offsets[i + 1];
      if (hi < lo) return MMB_ERR_ARG;
      encode_word(*b, words + lo, (size_t)(hi - lo), ids);
      out_counts[i] = (int)ids.size();
      for (int v : ids) out_ids[w++] = v;
    }
  } catch (const std::exception&) {
    return MMB_ERR_ARG;
  }
  return MMB_OK;
}

extern "C" int mmb_bpe_token_id(void* handle, const char* token) {
  if (!handle || !token) return -1;
  Bpe* b = static_cast<Bpe*>(handle);
  auto it = b->vocab.find(token);
  return it == b->vocab.end() ? -1 : it->second;
}

extern "C" int mmb_bpe_destroy(void* handle) {
  delete static_cast<Bpe*>(handle);
  return MMB_OK;
}
