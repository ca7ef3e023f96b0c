// b2enc — stand-alone .bz2 writer over the C++ mirror, the counterpart of the reference's
// extras/bzip2_enc.adb (:15-56): unbuffered byte callbacks, no size hint.
//   b2enc <in> <out> [-1|-2|-3]        (-1/-2/-3 = block_100k/400k/900k, as bzip2_enc.adb:82-84)
// Build:  g++ -O2 -std=c++17 b2enc.cpp -L.. -lb2gpu -Wl,-rpath,'$ORIGIN/..' -o b2enc
#include <cstdio>
#include <cstring>

#include "bzip2_encoding.hpp"

int main(int argc, char **argv) {
  if (argc < 3) { std::fprintf(stderr, "usage: b2enc <in> <out> [-1|-2|-3]\n"); return 2; }
  bzip2_encoding::Compression_Option opt = bzip2_encoding::block_900k;
  if (argc > 3 && !std::strcmp(argv[3], "-1")) opt = bzip2_encoding::block_100k;
  if (argc > 3 && !std::strcmp(argv[3], "-2")) opt = bzip2_encoding::block_400k;
  FILE *fi = std::fopen(argv[1], "rb"), *fo = std::fopen(argv[2], "wb");
  if (!fi || !fo) { std::perror("open"); return 1; }
  int nextc = std::fgetc(fi);
  try {
    bzip2_encoding::Encode([&] { uint8_t b = (uint8_t)nextc; nextc = std::fgetc(fi); return b; },
                           [&] { return nextc != EOF; },
                           [&](uint8_t b) { std::fputc(b, fo); }, opt);
  } catch (const std::exception &e) { std::fprintf(stderr, "b2enc: %s\n", e.what()); return 1; }
  std::fclose(fi); std::fclose(fo);
  return 0;
}
