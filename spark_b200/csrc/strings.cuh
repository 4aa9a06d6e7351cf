// strings.cuh -- string key columns as order-preserving int32 dictionary codes (csrc/strings.cu)
#pragma once
#include <vector>
#include "common.cuh"

namespace sb {

// codes: int32 column sharing `strings`' validity; dictionary: the distinct non-NULL values in UTF8String order
void dictionary_encode(const Column &strings, cudaStream_t st, Column &codes, Column &dictionary);
// codes of `strings` in an existing dictionary (-1: not in it; NULL stays NULL)
Column dictionary_lookup(const Column &strings, const Column &dictionary, cudaStream_t st);
// strings for codes (NULL / out-of-dictionary codes -> NULL)
Column dictionary_decode(const Column &codes, const Column &dictionary, cudaStream_t st);

// A zero-copy view of a table in which some string columns are replaced by their codes.  Owns the view and the dictionaries.
struct EncodedView {
  sb_table *view = nullptr;
  std::vector<int> cols;              // which columns were replaced
  std::vector<Column> dictionaries;   // one per replaced column
  EncodedView() = default;
  EncodedView(const EncodedView &) = delete;
  EncodedView &operator=(const EncodedView &) = delete;
  ~EncodedView();
  const Column *dictionary_of(int col) const {
    for (size_t i = 0; i < cols.size(); i++)
      if (cols[i] == col) return &dictionaries[i];
    return nullptr;
  }
};
// cols: string columns of t to encode (duplicates allowed).  dictionaries == nullptr: build them; otherwise look the values up
// in the given ones (same order as cols).
void encode_string_columns(const sb_table *t, const std::vector<int> &cols, const std::vector<const Column *> *dictionaries, cudaStream_t st,
                           EncodedView &out);

}  // namespace sb
