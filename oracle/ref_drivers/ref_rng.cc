// ref_rng: print raw draws of the reference's RNG layer so the oracle's restatement of
// MotherOfAll (newran1.cxx:383-432) and Normal (newran2.cxx:164-216,297-311) can be pinned.
// OUR code, compiled against the reference sources in place by oracle/build_ref.sh.
//   ref_rng <seed in (0,1)> <n>   -> n lines "u" from MotherOfAll(seed), then a second
//   generator MotherOfAll(seed) feeding n GaussianDist(0,1).draw() values (ProbabilityDist.cxx:64-89).
#include <cstdio>
#include <cstdlib>
#include <memory>
#include "newran.h"
#include "ProbabilityDist.h"
using namespace std;
int main(int argc, char **argv) {
  double seed = argc > 1 ? atof(argv[1]) : 0.224;
  int n = argc > 2 ? atoi(argv[2]) : 16;
  MotherOfAll a(seed);
  for (int i = 0; i < n; i++) printf("%.17g\n", a.Next());
  MotherOfAll b(seed);
  GaussianDist g(0.0, 1.0);
  for (int i = 0; i < n; i++) printf("%.17g\n", g.draw(&b));
  // master-seed derivation used by chain::chain() (chain.hh:58-59)
  ProbabilityDist::setSeed(seed);
  for (int i = 0; i < 4; i++) {
    MotherOfAll c(ProbabilityDist::getPRNG()->Next());
    printf("%.17g\n", c.Next());
  }
  return 0;
}
