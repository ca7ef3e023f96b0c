// C++ mirror of the reference's generic `BZip2.Encoding.Encode` (zip_lib/bzip2-encoding.ads:47-56)
// above the C ABI of include/b2gpu.h.  Same shape as the Ada generic: three callbacks for the
// bytes, `option` and `size_hint` as arguments; exceptions thrown by the callbacks propagate and
// the handle is released on every path (cf. bzip2-encoding.adb:1377-1381).  No CPU fallback.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/b2gpu.h"

namespace bzip2_encoding {

enum Compression_Option { block_100k = B2_BLOCK_100K, block_400k = B2_BLOCK_400K, block_900k = B2_BLOCK_900K };
using Stream_Size_Type = int64_t;
constexpr Stream_Size_Type unknown_size = B2_UNKNOWN_SIZE;

struct b2gpu_error : std::runtime_error { using std::runtime_error::runtime_error; };

// generic
//   with function Read_Byte return Byte; with function More_Bytes return Boolean;
//   with procedure Write_Byte (b : Byte);
// procedure Encode (option := block_900k; size_hint := unknown_size);
template <class ReadByte, class MoreBytes, class WriteByte>
void Encode(ReadByte Read_Byte, MoreBytes More_Bytes, WriteByte Write_Byte,
            Compression_Option option = block_900k, Stream_Size_Type size_hint = unknown_size, int device = 0) {
  std::vector<uint8_t> in;
  while (More_Bytes()) in.push_back(Read_Byte());
  b2_encoder *enc = nullptr;
  if (b2_create((int)option, device, &enc) != B2_OK) throw b2gpu_error(std::string("b2_create: ") + b2_last_error());
  struct Guard { b2_encoder *e; ~Guard() { b2_destroy(e); } } guard{enc};
  std::vector<uint8_t> out(b2_bound(in.size()) + 1024 * (in.size() / 40000 + 16));
  uint64_t out_len = 0;
  if (b2_encode_stream(enc, in.data(), in.size(), size_hint, out.data(), out.size(), &out_len) != B2_OK)
    throw b2gpu_error(std::string("b2_encode_stream: ") + b2_last_error());
  for (uint64_t i = 0; i < out_len; i++) Write_Byte(out[i]);
}

}  // namespace bzip2_encoding
