/* TEST INFRASTRUCTURE — plain-C restatement of the codebook nearest-neighbour search.
 *
 * Follows viewformer/models/utils_th.py:34-41 (QuantizeEMA.forward):
 *     dist = |z|^2 - 2 z.E + |E|^2          (fp32, codebook stored [D,K])
 *     _, ind = (-dist).max(1)               (first maximum wins on exact ties)
 * and utils_th.py:66  diff = mean((quantize - z)^2).
 * Built by oracle/Makefile into oracle/_ref/libvq_oracle.so; used only by tests/ and by the
 * cpu_baseline leg of bench.py.  Never part of the shipped CUDA path.
 */
#include <stdint.h>
#include <stddef.h>

/* z [M,D] row-major, E [D,K] row-major (reference layout), idx out [M]. */
void vq_oracle_lookup(const float *z, const float *E, int64_t M, int D, int K, int64_t *idx, double *diff_sum)
{
    double dsum = 0.0;
    for (int64_t m = 0; m < M; ++m) {
        const float *zr = z + (size_t)m * D;
        float zz = 0.f;
        for (int d = 0; d < D; ++d) zz += zr[d] * zr[d];
        float best = 0.f;
        int64_t bi = -1;
        for (int k = 0; k < K; ++k) {
            float dot = 0.f, ee = 0.f;
            for (int d = 0; d < D; ++d) {
                float e = E[(size_t)d * K + k];
                dot += zr[d] * e;
                ee += e * e;
            }
            float neg = -((zz - 2.f * dot) + ee);
            if (bi < 0 || neg > best) { best = neg; bi = k; }
        }
        idx[m] = bi;
        if (diff_sum) {
            for (int d = 0; d < D; ++d) {
                double t = (double)E[(size_t)d * K + bi] - (double)zr[d];
                dsum += t * t;
            }
        }
    }
    if (diff_sum) *diff_sum = dsum;
}
