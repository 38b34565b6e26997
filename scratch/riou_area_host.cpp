// host check of csrc/riou_area.cuh (same arithmetic in fp32 on the CPU) against the float64 oracle:
//   g++ -O2 -ffp-contract=off scratch/riou_area_host.cpp -Ioracle -Lor... (see scratch/riou_area_host.sh)
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include "../rotate-yolov3_b200/csrc/riou_area.cuh"
extern "C" double orc_skew_iou(const double* box1, const double* box2, int mode);
static double urand() { return rand() / (RAND_MAX + 1.0); }
static void gen(double* b, double canvas, int kind) {
  b[0] = urand() * canvas; b[1] = urand() * canvas;
  double area = 792 + urand() * (15803 - 792), ratio = 4 + urand() * 5;
  b[2] = sqrt(area * ratio); b[3] = sqrt(area / ratio);
  b[4] = (urand() - 0.5) * M_PI;
  if (kind == 1) { b[4] = 0; b[0] = floor(b[0]); b[1] = floor(b[1]); b[2] = floor(b[2]); b[3] = floor(b[3]) + 1; }
  if (kind == 2) { b[4] = (rand() % 4) * M_PI / 2 - M_PI / 2; }
}
static float iou32(const double* A, const double* B) {
  float a[5], b[5];
  for (int i = 0; i < 5; i++) { a[i] = (float)A[i]; b[i] = (float)B[i]; }
  float inter = ryolo::clamp_integral_area2(a[0], a[1], cosf(a[4]), sinf(a[4]), 0.5f * a[2], 0.5f * a[3], b[0], b[1], cosf(b[4]),
                                            sinf(b[4]), 0.5f * b[2], 0.5f * b[3]);
  float aa = a[2] * a[3], ab = b[2] * b[3];
  inter = fminf(inter, fminf(aa, ab));
  float u = aa + ab - inter;
  return u == 0.f ? 0.f : inter / u;
}
int main() {
  srand(7);
  for (int kind = 0; kind < 4; kind++) {
    double worst_rel = 0, worst_abs = 0; long nz = 0, bad = 0;
    for (long it = 0; it < 2000000; it++) {
      double a[5], b[5];
      gen(a, kind == 3 ? 100.0 : 1024.0, kind == 3 ? 0 : kind); gen(b, kind == 3 ? 100.0 : 1024.0, kind == 3 ? 0 : kind);
      if (kind == 1 && (it & 3) == 0) { b[1] = a[1]; b[3] = a[3]; }      // collinear top/bottom edges
      if (kind == 1 && (it & 7) == 1) { for (int i = 0; i < 5; i++) b[i] = a[i]; }
      if (kind == 3) { for (int i = 0; i < 5; i++) b[i] = a[i] + (urand() - 0.5) * 1e-3 * (it % 1000); }   // near-identical
      float fa[5], fb[5]; double da[5], db[5];
      for (int i = 0; i < 5; i++) { fa[i] = (float)a[i]; fb[i] = (float)b[i]; da[i] = fa[i]; db[i] = fb[i]; }
      double ref = orc_skew_iou(da, db, 0);
      double got = iou32(da, db);
      double err = fabs(got - ref);
      if (ref > 0) nz++;
      if (err > 1e-4 * fabs(ref) + 1e-6) { bad++; if (bad < 5) printf("  kind %d bad: ref %.9g got %.9g  a=(%g %g %g %g %g) b=(%g %g %g %g %g)\n", kind, ref, got, a[0],a[1],a[2],a[3],a[4],b[0],b[1],b[2],b[3],b[4]); }
      if (err > worst_abs) worst_abs = err;
      if (ref > 1e-3 && err / ref > worst_rel) worst_rel = err / ref;
    }
    printf("kind %d: nonzero %ld, worst abs %.3g, worst rel (ref>1e-3) %.3g, out of tolerance %ld\n", kind, nz, worst_abs, worst_rel, bad);
  }
  return 0;
}
