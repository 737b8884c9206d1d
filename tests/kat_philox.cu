// Host-only known-answer check of philox4x32_10() in bwgr_b200/csrc/common.cuh (built and run by tests/test_blocked_math.py).
#include <cstdio>
#include "common.cuh"
int main() {
  uint32_t c[3][4] = {{0, 0, 0, 0}, {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu}, {0x243f6a88u, 0x85a308d3u, 0x13198a2eu, 0x03707344u}};
  uint32_t k[3][2] = {{0, 0}, {0xffffffffu, 0xffffffffu}, {0xa4093822u, 0x299f31d0u}};
  for (int i = 0; i < 3; i++) {
    bwgr::philox4x32_10(c[i], k[i][0], k[i][1]);
    printf("%08x %08x %08x %08x\n", c[i][0], c[i][1], c[i][2], c[i][3]);
  }
  return 0;
}
