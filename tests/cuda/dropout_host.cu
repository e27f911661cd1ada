// Host-side evaluation of the kernels' dropout keep function (csrc/mlt_common.cuh is __host__ __device__):
// prints keep(b, h, i, col) for a small grid so that tests/test_dropout_ref.py can pin the numpy
// restatement (tests/dropout_ref.py) to the very code the kernels compile.
#include <cstdio>
#include <cstdlib>

#include "../../multimodal-long-transformer-2021_b200/csrc/mlt_common.cuh"

int main(int argc, char** argv) {
  if (argc < 8) return 2;
  const unsigned long long seed = strtoull(argv[1], nullptr, 10);
  const double p = atof(argv[2]);
  const int B = atoi(argv[3]), H = atoi(argv[4]), rows = atoi(argv[5]), cols = atoi(argv[6]), rowset = atoi(argv[7]);
  mlt::Dropout d{};
  double t = p * 4294967296.0;
  if (t > 4294967295.0) t = 4294967295.0;
  d.thr = (uint32_t)t;
  d.seed_lo = (uint32_t)seed;
  d.seed_hi = (uint32_t)(seed >> 32);
  d.rowset = (uint32_t)rowset;
  for (int b = 0; b < B; ++b)
    for (int i = 0; i < rows; ++i)
      for (int c = 0; c < cols; ++c)
        for (int h = 0; h < H; ++h) {
          const uint32_t base = mlt::dropout_row_base(mlt::dropout_salt(d, (uint32_t)(b * H + h)), i);
          putchar(mlt::dropout_keep(base, c, d.thr) ? '1' : '0');
        }
  putchar('\n');
  return 0;
}
