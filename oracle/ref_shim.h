/*
 * oracle/ref_shim.h — TEST INFRASTRUCTURE.  Host-side stand-ins for the handful of CUDA
 * built-ins the reference SoftRas kernel source uses, so that the reference's OWN kernel
 * bodies (third-party/softras/soft_renderer/cuda/soft_rasterize_cuda_kernel.cu:22-671, the
 * anonymous namespace holding the three __global__ templates and their helpers) can be
 * compiled with g++ and executed on the CPU as the ground truth that pins oracle/softras_oracle.c.
 * No reference source is copied into this repository: oracle/Makefile extracts the namespace
 * block from /root/reference at build time into oracle/_ref/ (git-ignored).
 *
 * Overload resolution below reproduces CUDA's device-side math overloads:
 *   max/min(float,float) -> float ; any double operand -> double ;
 *   exp/sqrt/pow on float -> float versions (libstdc++ <cmath> already does that).
 */
#pragma once
#include <cmath>
#include <math.h>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline

struct scp_ref_dim3 { unsigned x, y, z; };
static thread_local scp_ref_dim3 blockIdx, threadIdx, blockDim;

static inline float  max(float a, float b)   { return fmaxf(a, b); }
static inline double max(float a, double b)  { return fmax((double)a, b); }
static inline double max(double a, float b)  { return fmax(a, (double)b); }
static inline double max(double a, double b) { return fmax(a, b); }
static inline float  min(float a, float b)   { return fminf(a, b); }
static inline double min(float a, double b)  { return fmin((double)a, b); }
static inline double min(double a, float b)  { return fmin(a, (double)b); }
static inline double min(double a, double b) { return fmin(a, b); }
static inline float  pow(float a, int b)     { return powf(a, (float)b); }

/* sequential per image in the driver, so a plain add is the atomic */
static inline float atomicAdd(float *p, float v) { float o = *p; *p = o + v; return o; }
static inline double atomicAdd(double *p, double v) { double o = *p; *p = o + v; return o; }
