/* newran_port.h -- TEST INFRASTRUCTURE (oracle).  Not part of the product; never linked by it.
 *
 * CPU restatement of the two generators of the vendored newran03 library that the reference's
 * hot path consumes, so that the oracle can replay a reference run draw for draw:
 *   MotherOfAll  (newran1.cxx:334-432): Marsaglia's two multiply-with-carry generators on
 *                SIGNED 16-bit lanes, uniform = (seed + 0.5) / 2^32
 *   Normal       (newran2.cxx:164-185 PosGen::Build, :202-216 SymGen::Next, :297-311 Normal):
 *                table-rejection sampler, 3 uniforms per attempt
 * Pinned against the real library through oracle/_ref/ref_rng (tests/test_oracle_ref.py).
 */
#ifndef NEWRAN_PORT_H
#define NEWRAN_PORT_H
#include <stdint.h>
typedef struct {
  int16_t m1[10], m2[10];
  int started;
  uint64_t seed;
} mother_t;
void mother_init(mother_t *m, double s);   /* MotherOfAll::MotherOfAll(double) */
double mother_next(mother_t *m);           /* MotherOfAll::Next()              */
double newran_normal(mother_t *m);         /* Normal::Next() drawing from m    */
#endif
