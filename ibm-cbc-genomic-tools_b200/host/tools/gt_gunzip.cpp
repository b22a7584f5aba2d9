// gt_gunzip -- what gt::GzipStream (host/gt_inflate.h) makes of a gzip file, on standard output: the tests hold it against zlib
// (whole files, truncated and damaged ones, several members, every block type).  Exit code 1: the stream ended on a fault.
//
//   gt_gunzip FILE [READ_SIZE]
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <unistd.h>
#include <vector>
#include "../gt_inflate.h"

int main(int argc, char **argv) {
  if (argc < 2) { fprintf(stderr, "usage: gt_gunzip FILE [READ_SIZE]\n"); return 2; }
  const int fd = open(argv[1], O_RDONLY);
  if (fd < 0) { fprintf(stderr, "gt_gunzip: cannot open '%s'\n", argv[1]); return 2; }
  gt::GzipStream gz(fd);
  std::vector<char> buf(argc > 2 ? (size_t)atol(argv[2]) : (size_t)3 << 20);
  for (;;) {
    const long n = gz.Read(buf.data(), buf.size());
    if (n <= 0) break;
    fwrite(buf.data(), 1, (size_t)n, stdout);
  }
  close(fd);
  return gz.failed() ? 1 : 0;
}
