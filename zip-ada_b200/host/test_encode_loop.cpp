// The call sequence of the Ada replacement body (zip-ada_b200/ada/bzip2-encoding.adb) as an unmodified
// Zip.Create.Add_Stream loop produces it: Encode once per archive entry (zip-create.adb:253-265), exceptions
// raised from the callbacks in the middle of an entry (Compression_inefficient from Write_Byte,
// zip-compress.adb:480-486; User_abort from Read_Byte, zip-compress-bzip2_e.adb:78-96), then more entries.
// Usage: test_encode_loop <out_dir> <in_file>...   writes <out_dir>/<k>.bz2 for every entry that completed and
// prints "created=<handles created> ok=<entries> inefficient=<k> aborted=<k>".
#include <cstdio>
#include <fstream>
#include <iterator>
#include <string>
#include <vector>

#include "bzip2_encoding.hpp"

struct Compression_inefficient {};
struct User_abort {};

int main(int argc, char **argv) {
  if (argc < 3) return 2;
  const std::string out_dir = argv[1];
  int ok = 0, ineff = 0, aborted = 0;
  for (int k = 2; k < argc; k++) {
    std::ifstream f(argv[k], std::ios::binary);
    std::vector<uint8_t> in((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    std::vector<uint8_t> out;
    size_t pos = 0;
    const int entry = k - 2;
    try {
      bzip2_encoding::Encode(
          [&]() -> uint8_t {
            if (entry == 2 && pos == in.size() / 2) throw User_abort();       // the user cancels in the middle of entry 2
            return in[pos++];
          },
          [&]() { return pos < in.size(); },
          [&](uint8_t b) {
            out.push_back(b);
            if (entry == 1 && out.size() >= in.size()) throw Compression_inefficient();   // output_size >= input_size
          },
          bzip2_encoding::block_900k, (int64_t)in.size());
      std::ofstream o(out_dir + "/" + std::to_string(entry) + ".bz2", std::ios::binary);
      o.write((const char *)out.data(), (std::streamsize)out.size());
      ok++;
    } catch (const Compression_inefficient &) {
      ineff++;
    } catch (const User_abort &) {
      aborted++;
    }
  }
  std::printf("created=%llu ok=%d inefficient=%d aborted=%d\n", (unsigned long long)bzip2_encoding::Handle_Pool::instance().created(), ok,
              ineff, aborted);
  return 0;
}
