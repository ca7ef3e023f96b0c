// Minimal counterpart of `zipada -eb3 archive.zip files...` (tools/zipada.adb) for the BZip2 methods:
// b2zip [-eb1|-eb2|-eb3] archive.zip file...
#include <cstdio>
#include <cstring>
#include <string>

#include "zip_create.hpp"

int main(int argc, char **argv) {
  using namespace zip_create;
  Compression_Method m = BZip2_3;
  int a = 1;
  if (a < argc && !strncmp(argv[a], "-eb", 3)) { m = argv[a][3] == '1' ? BZip2_1 : argv[a][3] == '2' ? BZip2_2 : BZip2_3; a++; }
  if (argc - a < 1) { fprintf(stderr, "usage: b2zip [-eb1|-eb2|-eb3] archive.zip file...\n"); return 2; }
  try {
    Zip_Create_Info info;
    info.Create_Archive(argv[a++], m);
    for (; a < argc; a++) info.Add_File(argv[a]);
    info.Finish();
    for (const auto &e : info.Entries()) printf("method %2u  %llu bytes\n", e.zip_type, (unsigned long long)e.compressed_size);
  } catch (const std::exception &ex) {
    fprintf(stderr, "b2zip: %s\n", ex.what());
    return 1;
  }
  return 0;
}
