/* newran_port.c -- TEST INFRASTRUCTURE (oracle); see newran_port.h.  Never linked by the product. */
#include "newran_port.h"
#include <math.h>
#include <string.h>

/* MotherOfAll::MotherOfAll(double s), newran1.cxx:334-344 */
void mother_init(mother_t *m, double s) {
  memset(m, 0, sizeof(*m));
  m->seed = (uint64_t)(s * 2147483648.0);
  m->started = 0;
}

/* MotherOfAll::Mother(), newran1.cxx:383-430.  NB the lanes are SIGNED shorts in the reference. */
static void mother_step(mother_t *m) {
  uint64_t number, number1, number2;
  if (!m->started) {
    uint16_t sNumber = (uint16_t)(m->seed & 0xFFFF);
    number = m->seed & 0x7FFFFFFF;
    int16_t *p = m->m1;
    for (int n = 17; n >= 0; n--) {
      number = (uint64_t)(int64_t)(30903 * (int)sNumber) + (number >> 16);
      sNumber = (uint16_t)(number & 0xFFFF);
      *p++ = (int16_t)sNumber;
      if (n == 9) p = m->m2;
    }
    m->m1[0] &= 0x7FFF; m->m2[0] &= 0x7FFF;
    m->started = 1;
  }
  memmove(m->m1 + 2, m->m1 + 1, 8 * sizeof(int16_t));
  memmove(m->m2 + 2, m->m2 + 1, 8 * sizeof(int16_t));
  number1 = (uint64_t)(int64_t)m->m1[0];
  number2 = (uint64_t)(int64_t)m->m2[0];
  {
    int a = 1941 * m->m1[2] + 1860 * m->m1[3] + 1812 * m->m1[4] + 1776 * m->m1[5] +
            1492 * m->m1[6] + 1215 * m->m1[7] + 1066 * m->m1[8] + 12013 * m->m1[9];
    int b = 1111 * m->m2[2] + 2222 * m->m2[3] + 3333 * m->m2[4] + 4444 * m->m2[5] +
            5555 * m->m2[6] + 6666 * m->m2[7] + 7777 * m->m2[8] + 9272 * m->m2[9];
    number1 += (uint64_t)(int64_t)a;
    number2 += (uint64_t)(int64_t)b;
  }
  m->m1[0] = (int16_t)(uint16_t)(number1 / 65536u);
  m->m2[0] = (int16_t)(uint16_t)(number2 / 65536u);
  m->m1[1] = (int16_t)(uint16_t)(0xFFFF & number1);
  m->m2[1] = (int16_t)(uint16_t)(0xFFFF & number2);
  {
    int64_t t = ((int64_t)m->m1[1]) * 65536 + (int64_t)m->m2[1];
    m->seed = ((uint64_t)t) & 0xFFFFFFFFu;
  }
}

/* MotherOfAll::Next(), newran1.cxx:432 */
double mother_next(mother_t *m) {
  mother_step(m);
  return ((double)m->seed + 0.5) / 4294967296.0;
}

/* Normal::Density, newran2.cxx:310-311 */
static double normal_density(double x) { return (fabs(x) > 8.0) ? 0 : 0.398942280 * exp(-x * x / 2); }

static double n_sx[60], n_sfx[60], n_xi;
static int n_built = 0;
/* PosGen::Build(true), newran2.cxx:164-185 */
static void normal_build(void) {
  double sxi = 0.0, inc = 0.01;
  int i;
  for (i = 0; i < 60; i++) {
    n_sx[i] = sxi;
    double f1 = normal_density(sxi);
    n_sfx[i] = f1;
    if (f1 <= 0.0) break;
    sxi += inc / f1;
  }
  n_xi = 2 * i;
  n_built = 1;
}

/* SymGen::Next(), newran2.cxx:202-216 */
double newran_normal(mother_t *m) {
  double s, ak, y;
  int ir;
  if (!n_built) normal_build();
  do {
    s = 1.0;
    double r1 = mother_next(m);
    if (r1 > 0.5) { s = -1.0; r1 = 1.0 - r1; }
    ir = (int)(r1 * n_xi);
    double sxi = n_sx[ir];
    ak = sxi + (n_sx[ir + 1] - sxi) * mother_next(m);
    y = n_sfx[ir] * mother_next(m);
  } while (y >= n_sfx[ir + 1] && y >= normal_density(ak));
  return s * ak;
}
