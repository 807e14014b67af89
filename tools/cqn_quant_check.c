/* cqn_quant_check.c — exhaustive check of the quantiser shortcut of vorbis_b200/csrc/vb200_cqn.cuh (cqn_quant):
 * for EVERY float ve in [0, 2^22) the value q0 + (ve > q0^2+q0+.25 || (ve == .. && q0 odd)), q0 = (int)sqrtf(ve),
 * must equal (int)rint(sqrt((double)ve)) - the expression of the reference (lib/psy.c:959-963).  sqrtf here is the
 * correctly rounded IEEE square root, as on the device (-prec-sqrt=true).  About 1.25e9 values; OpenMP optional.
 *
 *   gcc -O2 -fopenmp -ffp-contract=off tools/cqn_quant_check.c -lm -o /tmp/cqn_quant_check && /tmp/cqn_quant_check [stride]
 * stride > 1 checks every stride-th bit pattern (the CPU test uses a stride, the full run is this tool's default). */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static int shortcut(float ve){
  const int q0 = (int)sqrtf(ve);
  const double h = (double)(q0*q0 + q0) + .25, vd = (double)ve;
  return q0 + ((vd > h || (vd == h && (q0 & 1))) ? 1 : 0);
}

int main(int argc, char **argv){
  const uint32_t stride = argc > 1 ? (uint32_t)strtoul(argv[1], 0, 10) : 1;
  uint32_t top; float lim = 4194304.f;
  long long bad = 0, n = 0;
  memcpy(&top, &lim, 4);                       /* bit pattern of 2^22: every smaller non-negative float precedes it */
#pragma omp parallel for reduction(+:bad,n) schedule(static)
  for(long long u = 0; u < (long long)top; u += stride){
    uint32_t bits = (uint32_t)u; float ve; memcpy(&ve, &bits, 4);
    const int want = (int)rint(sqrt((double)ve));
    if(shortcut(ve) != want){ bad++; if(bad < 5) fprintf(stderr, "mismatch at %.9g: %d vs %d\n", ve, shortcut(ve), want); }
    n++;
  }
  printf("{\"checked\": %lld, \"stride\": %u, \"mismatches\": %lld}\n", n, stride, bad);
  return bad != 0;
}
