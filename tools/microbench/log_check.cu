// Accuracy and cost of the Kennedy-Pendleton log (update.cuh, Mth<double>::log_) against libdevice's log on x = 1 - r, r a float in (0, 1).
#include <cstdio>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../../qcdgpu_b200/csrc/launch.h"
#include "../../qcdgpu_b200/csrc/update.cuh"
using namespace qcd;
template <int WHICH> __global__ void k(const float* r, double* out, int n, int reps) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x = 1.0 - double(r[i]), acc = 0.0;
    for (int j = 0; j < reps; j++) { acc += WHICH ? log(x) : Mth<double>::log_(x); x = fma(x, 0.999999, 1e-9); }
    out[i] = acc;
}
__global__ void ktrig(const float* r, double* out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s, c;
    Mth<double>::sincos2pi(double(r[i]), &s, &c);
    out[3 * i] = s; out[3 * i + 1] = c; out[3 * i + 2] = Mth<double>::cos2pi(double(r[i]));
}
int main() {
    const int n = 1 << 22;
    std::vector<float> h(n);
    srand(3);
    for (int i = 0; i < n; i++) {
        uint32_t m = ((uint32_t)rand() & 0xffffff) | 1u;
        h[i] = (i % 7 == 0) ? ldexpf(float(m), -24 - (rand() % 20)) : (i % 11 == 0) ? 1.0f - ldexpf(float((rand() & 0xfff) + 1), -24) : ldexpf(float(m), -24);
        if (!(h[i] > 0.0f && h[i] < 1.0f)) h[i] = 0.5f;
    }
    float* dr; double *da, *db;
    cudaMalloc(&dr, n * 4); cudaMalloc(&da, n * 8); cudaMalloc(&db, n * 8);
    cudaMemcpy(dr, h.data(), n * 4, cudaMemcpyHostToDevice);
    k<0><<<n / 256, 256>>>(dr, da, n, 1); k<1><<<n / 256, 256>>>(dr, db, n, 1);
    std::vector<double> a(n), b(n);
    cudaMemcpy(a.data(), da, n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(b.data(), db, n * 8, cudaMemcpyDeviceToHost);
    double maxulp = 0, maxabs = 0; int worst = 0;
    for (int i = 0; i < n; i++) {
        double want = (double)logl(1.0L - (long double)h[i]);
        double ulp = fabs(a[i] - want) / (fabs(want) * 2.220446049250313e-16 + 1e-300);
        if (want != 0.0 && ulp > maxulp) { maxulp = ulp; worst = i; }
        if (fabs(a[i] - b[i]) > maxabs) maxabs = fabs(a[i] - b[i]);
    }
    printf("max error of log_ vs long-double log: %.3f ulp (r = %.9g), max |log_ - libdevice log| = %.3e\n", maxulp, h[worst], maxabs);
    {   // sin / cos of 2 pi x against long double
        double* dt; cudaMalloc(&dt, size_t(n) * 24);
        ktrig<<<n / 256, 256>>>(dr, dt, n);
        std::vector<double> t(size_t(n) * 3);
        cudaMemcpy(t.data(), dt, size_t(n) * 24, cudaMemcpyDeviceToHost);
        double worst = 0;
        for (int i = 0; i < n; i++) {
            long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)h[i];
            double es = fabs(t[3 * i] - (double)sinl(a)), ec = fabs(t[3 * i + 1] - (double)cosl(a)), e2 = fabs(t[3 * i + 2] - (double)cosl(a));
            worst = fmax(worst, fmax(es, fmax(ec, e2)));
        }
        printf("max absolute error of sincos2pi / cos2pi vs long double: %.3e\n", worst);
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 2; w++) {
        cudaEventRecord(e0);
        if (w == 0) k<0><<<n / 256, 256>>>(dr, da, n, 64); else k<1><<<n / 256, 256>>>(dr, db, n, 64);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%s: %.3f ms for %d x 64 logs\n", w == 0 ? "log_ (this)" : "log (libdevice)", ms, n);
    }
    return 0;
}
