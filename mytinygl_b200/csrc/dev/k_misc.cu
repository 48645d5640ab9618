/*
 * k_misc.cu -- K6 and small utility kernels.
 *
 *   k_mip1        texture_generate_mip1 (src/textures.c:311-354): 2x2 box filter, sum of four
 *                 n/255 floats * 0.25, truncating pack; built eagerly at upload instead of lazily
 *                 inside the sampler (identical values).
 *   k_fill_unorm8 the 256 correctly rounded quotients n / 255.0f that color_from_rgba32
 *                 (src/graphics.h:350-357) produces, computed with the IEEE division itself.
 */
#include "dev_common.cuh"

namespace mtgl_dev_impl {

void note_launch();

__global__ void k_fill_unorm8(float *table)
{
    int i = threadIdx.x;
    if (i < 256) table[i] = (float)i / 255.0f;
}

__global__ void __launch_bounds__(256) k_mip1(const uint32_t *l0, int w, int h, uint32_t *l1, const float *unorm8)
{
    const int w1 = w / 2, h1 = h / 2;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w1 * h1) return;
    int x = i % w1, y = i / w1;
    Color4 a = color_unpack(l0[(2 * y) * w + 2 * x], unorm8);
    Color4 b = color_unpack(l0[(2 * y) * w + 2 * x + 1], unorm8);
    Color4 c = color_unpack(l0[(2 * y + 1) * w + 2 * x], unorm8);
    Color4 d = color_unpack(l0[(2 * y + 1) * w + 2 * x + 1], unorm8);
    Color4 s = { ((a.r + b.r) + c.r) + d.r, ((a.g + b.g) + c.g) + d.g, ((a.b + b.b) + c.b) + d.b, ((a.a + b.a) + c.a) + d.a };
    s.r *= 0.25f; s.g *= 0.25f; s.b *= 0.25f; s.a *= 0.25f;
    l1[i] = color_pack(s);
}

/* float4 copy of a small texture (level 0, then level 1): every texel as color_from_rgba32 (src/graphics.h:350-357) returns
 * it, so that a tap of the tile kernels' samplers (dev_fasttex.cuh) is one 16-byte load with no conversion */
__global__ void __launch_bounds__(256) k_tex_f4(const uint32_t *l0, int n0, const uint32_t *l1, int n1, float4 *out, const float *unorm8)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n0 + n1) return;
    const Color4 c = color_unpack(i < n0 ? l0[i] : l1[i - n0], unorm8);
    out[i] = make_float4(c.r, c.g, c.b, c.a);
}

void launch_tex_f4(const uint32_t *l0, int n0, const uint32_t *l1, int n1, float4 *out, const float *unorm8, cudaStream_t s)
{
    if (n0 + n1 <= 0) return;
    k_tex_f4<<<(n0 + n1 + 255) / 256, 256, 0, s>>>(l0, n0, l1, n1, out, unorm8);
    note_launch();
}

/* Small batch arenas (state blocks, draw records: a few KB) are pulled over PCIe by a kernel reading the pinned,
 * device-mapped staging buffer instead of a DMA: a host-to-device memcpy between two frames' kernels costs a
 * compute -> copy-engine -> compute round trip, several times the transfer itself. */
__global__ void __launch_bounds__(256) k_upload(const uint4 *__restrict__ src, uint4 *__restrict__ dst, uint32_t n16, uint4 *__restrict__ zero, uint32_t z16)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x) dst[i] = src[i];
    /* the tables a pass starts from zero (counters, per-tile counts and flags), in the same launch */
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < z16; i += gridDim.x * blockDim.x) zero[i] = make_uint4(0u, 0u, 0u, 0u);
}

/* zero / zero_bytes (a multiple of 16, may be 0): cleared by the same kernel */
void launch_upload(const void *host_mapped, void *dst, size_t bytes, void *zero, size_t zero_bytes, cudaStream_t s)
{
    const uint32_t n16 = (uint32_t)((bytes + 15) / 16), z16 = (uint32_t)(zero_bytes / 16);
    if (!n16 && !z16) return;
    k_upload<<<min((max(n16, z16) + 255u) / 256u, 148u * 4u), 256, 0, s>>>(static_cast<const uint4 *>(host_mapped), static_cast<uint4 *>(dst), n16,
                                                                           static_cast<uint4 *>(zero), z16);
    note_launch();
}

/* Frame barrier of a sort-first multi-GPU frame, in stream order behind a device's raster kernels: "my band has landed
 * in the presenting GPU's plane" (arrive: one system-scope atomic on a counter that lives in that plane's allocation,
 * over NVLink for every GPU but the presenter), then wait until all participants of this frame have arrived.  It
 * replaces an NCCL all-reduce used as a barrier (launch + protocol + two cross-stream event waits: 50-100 us per
 * frame on 8 GPUs) by one tiny kernel on the stream that is busy anyway.  The colour stores of the preceding kernels
 * are complete when this kernel starts (stream order); the fence orders them before the arrival for good measure.
 * A participant that never arrives would make the others spin for ever: after timeout_ns (MTGL_BARRIER_TIMEOUT_S, default
 * 10 s, 0 = wait for ever) the kernel gives up, says so in mapped host memory and ends normally -- the host reports
 * MTGL_E_CUDA from the next mtgl_dev_finish, the context stays usable (a trap would poison it on every rank). */
__global__ void k_frame_barrier(unsigned long long *counter, unsigned long long target, unsigned long long timeout_ns, uint32_t *timed_out)
{
    __threadfence_system();
    atomicAdd_system(counter, 1ull);
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (atomicAdd_system(counter, 0ull) < target) {
        __nanosleep(200);
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (timeout_ns && t - t0 > timeout_ns) { *timed_out = 1u; break; }
    }
    __threadfence_system();
}

void launch_frame_barrier(unsigned long long *counter, unsigned long long target, unsigned long long timeout_ns, uint32_t *timed_out, cudaStream_t s)
{
    k_frame_barrier<<<1, 1, 0, s>>>(counter, target, timeout_ns, timed_out);
    note_launch();
}

void launch_fill_unorm8(float *table, cudaStream_t s)
{
    k_fill_unorm8<<<1, 256, 0, s>>>(table);
    note_launch();
}

void launch_mip1(const uint32_t *l0, int w, int h, uint32_t *l1, const float *unorm8, cudaStream_t s)
{
    int n = (w / 2) * (h / 2);
    if (n <= 0) return;
    k_mip1<<<(n + 255) / 256, 256, 0, s>>>(l0, w, h, l1, unorm8);
    note_launch();
}

} // namespace mtgl_dev_impl
