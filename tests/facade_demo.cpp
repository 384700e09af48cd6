// A caller written the way the OpenMC fork uses the library: only
// <pumitally/PumiTally.h> and -lpumitally.  Replays move 1 of the reference's
// known-answer test (test/test_pumi_tally_impl_methods.cpp:176-282).
#include <cstdio>
#include <string>
#include <vector>

#include "pumitally/PumiTally.h"

int main(int argc, char **argv) {
  const int n = 5;
  pumitally::PumiTally tally("box:1,1,1", n, argc, argv);
  std::vector<double> init(3 * n), dest(3 * n), w(n, 1.0);
  std::vector<int8_t> flying(n, 1);
  for (int i = 0; i < n; ++i) {
    init[3 * i] = 0.1; init[3 * i + 1] = 0.4; init[3 * i + 2] = 0.5;
    dest[3 * i] = 1.2; dest[3 * i + 1] = 0.4; dest[3 * i + 2] = 0.5;
  }
  tally.CopyInitialPosition(init.data(), 3 * n);
  tally.MoveToNextLocation(init.data(), dest.data(), flying.data(), w.data(), 3 * n);
  for (int i = 0; i < n; ++i)
    if (flying[i] != 0) { printf("flying not reset\n"); return 1; }
  tally.WriteTallyResults();
  printf("FACADE_OK\n");
  return 0;
}
