// Host-side execution of the SAME philox4x32_10 / dropout_words source the kernels use (csrc/philox.cuh is __host__ __device__):
// prints the words for the counters given on the command line so tests/test_philox.py can compare them with oracle/philox.py.
#include <cstdio>
#include <cstdlib>
#include "../vl-bert_b200/csrc/philox.cuh"

int main(int argc, char** argv) {
  if (argc == 7) {   // c0 c1 c2 c3 k0 k1
    uint32_t v[6];
    for (int i = 0; i < 6; ++i) v[i] = (uint32_t)strtoul(argv[i + 1], nullptr, 0);
    const vlb::Philox4 r = vlb::philox4x32_10(v[0], v[1], v[2], v[3], v[4], v[5]);
    printf("%08x %08x %08x %08x\n", r.x, r.y, r.z, r.w);
    return 0;
  }
  if (argc == 6) {   // n p seed site step -> keep flags of n elements as a 0/1 string
    const long n = atol(argv[1]);
    const float p = (float)atof(argv[2]);
    const uint64_t seed = strtoull(argv[3], nullptr, 0);
    const uint32_t site = (uint32_t)atoi(argv[4]), step = (uint32_t)atoi(argv[5]);
    const uint32_t t = vlb::dropout_threshold(p);
    for (long g = 0; g * 4 < n; ++g) {
      const vlb::Philox4 r = vlb::dropout_words((uint64_t)g, seed, site, step);
      const uint32_t w[4] = {r.x, r.y, r.z, r.w};
      for (int j = 0; j < 4 && g * 4 + j < n; ++j) putchar(w[j] >= t ? '1' : '0');
    }
    putchar('\n');
    return 0;
  }
  return 2;
}
