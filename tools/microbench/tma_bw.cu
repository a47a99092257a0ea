// micro-benchmark: device copy bandwidth through (a) cp.async.bulk load+store per warp tile,
// (b) plain LDG.128/STG.128, (c) bulk loads only (data dropped), for different tile sizes / warps.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t sa(const void* p){ return (uint32_t)__cvta_generic_to_shared(p);}        
__device__ __forceinline__ void mbar_init(uint64_t* b, unsigned c){ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;"::"r"(sa(b)),"r"(c)); asm volatile("fence.mbarrier_init.release.cluster;":::"memory");}
__device__ __forceinline__ void mbar_expect(uint64_t* b, unsigned n){ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"::"r"(sa(b)),"r"(n):"memory");}
__device__ __forceinline__ void mbar_wait(uint64_t* b, unsigned ph){ asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra W_%=;\n}"::"r"(sa(b)),"r"(ph):"memory");}
__device__ __forceinline__ void g2s(void* d, const void* s, unsigned n, uint64_t* b){ asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"::"r"(sa(d)),"l"(s),"r"(n),"r"(sa(b)):"memory");}
__device__ __forceinline__ void s2g(void* d, const void* s, unsigned n){ asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"::"l"(d),"r"(sa(s)),"r"(n):"memory"); asm volatile("cp.async.bulk.commit_group;":::"memory");}
template<int N> __device__ __forceinline__ void swait(){ asm volatile("cp.async.bulk.wait_group.read %0;"::"n"(N):"memory");}
extern __shared__ __align__(128) uint8_t smem[];

template<int TILE, int STAGES, bool STORE>
__global__ void k_tma(const uint8_t* in, uint8_t* out, size_t ntiles, int warps)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint8_t* tiles = smem + (size_t)w * STAGES * TILE;
    uint64_t* bars = (uint64_t*)(smem + (size_t)warps * STAGES * TILE) + w * STAGES;
    if (lane == 0) for (int s = 0; s < STAGES; s++) mbar_init(&bars[s], 1);
    __syncwarp();
    const size_t nw = (size_t)gridDim.x * warps, w0 = (size_t)blockIdx.x * warps + w;
    unsigned ph = 0; int st = 0;
    if (lane == 0) for (int s = 0; s < STAGES - 1; s++) { size_t t = w0 + s * nw; if (t < ntiles) { mbar_expect(&bars[s], TILE); g2s(tiles + s * TILE, in + t * TILE, TILE, &bars[s]); } }
    uint32_t acc = 0;
    for (size_t t = w0; t < ntiles; t += nw) {
        if (lane == 0) {
            if (STORE) swait<STAGES - 2 >= 0 ? (STAGES >= 3 ? 1 : 0) : 0>();
            size_t tn = t + (size_t)(STAGES - 1) * nw; int sn = (st + STAGES - 1) % STAGES;
            if (tn < ntiles) { mbar_expect(&bars[sn], TILE); g2s(tiles + sn * TILE, in + tn * TILE, TILE, &bars[sn]); }
        }
        __syncwarp();
        mbar_wait(&bars[st], (ph >> st) & 1); ph ^= 1u << st;
        acc += tiles[st * TILE + lane * 16];
        if (STORE) { asm volatile("fence.proxy.async.shared::cta;":::"memory"); __syncwarp(); if (lane == 0) s2g(out + t * TILE, tiles + st * TILE, TILE); }
        st = (st + 1) % STAGES;
    }
    if (STORE && lane == 0) swait<0>();
    if (acc == 0xdeadbeef) out[0] = 1;
}

__global__ void k_ldg(const uint4* in, uint4* out, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, s = (size_t)gridDim.x * blockDim.x;
    for (; i + 3 * s < n; i += 4 * s) { uint4 a = in[i], b = in[i + s], c = in[i + 2 * s], d = in[i + 3 * s]; out[i] = a; out[i + s] = b; out[i + 2 * s] = c; out[i + 3 * s] = d; }
    for (; i < n; i += s) out[i] = in[i];
}

template<typename F> float timeit(F f){ cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b); f(); float best=1e9; for(int r=0;r<5;r++){ cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms,a,b); best = ms<best?ms:best;} return best; }

template<int TILE, int STAGES, bool STORE> void run(const uint8_t* in, uint8_t* out, size_t bytes, int warps)
{
    size_t ntiles = bytes / TILE; size_t sm = (size_t)warps * STAGES * TILE + warps * STAGES * 8 + 64;
    cudaFuncSetAttribute(k_tma<TILE,STAGES,STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    float ms = timeit([&]{ k_tma<TILE,STAGES,STORE><<<148, warps*32, sm>>>(in, out, ntiles, warps); });
    cudaError_t e = cudaGetLastError();
    printf("tma tile %6d stages %d warps %d store %d: %.3f ms  %.0f GB/s (%s)\n", TILE, STAGES, warps, (int)STORE, ms, (STORE?2.0:1.0)*ntiles*TILE/ms/1e6, cudaGetErrorString(e));
}

int main(){
    size_t bytes = (size_t)1728 << 20; uint8_t *in, *out; cudaMalloc(&in, bytes); cudaMalloc(&out, bytes); cudaMemset(in, 1, bytes); cudaMemset(out, 0, bytes);
    float ms = timeit([&]{ k_ldg<<<148*8, 256>>>((const uint4*)in, (uint4*)out, bytes/16); });
    printf("ldg/stg copy: %.3f ms %.0f GB/s\n", ms, 2.0*bytes/ms/1e6);
    ms = timeit([&]{ cudaMemcpyAsync(out, in, bytes, cudaMemcpyDeviceToDevice); });
    printf("cudaMemcpy D2D: %.3f ms %.0f GB/s\n", ms, 2.0*bytes/ms/1e6);
    run<13824,2,true>(in,out,bytes,8); run<13824,4,true>(in,out,bytes,4); run<13824,3,true>(in,out,bytes,5);
    run<13824,2,false>(in,out,bytes,8); run<13824,4,false>(in,out,bytes,4);
    run<4096,4,true>(in,out,bytes,12); run<4096,4,false>(in,out,bytes,12); run<32768,3,true>(in,out,bytes,2); run<32768,3,false>(in,out,bytes,2);
    run<1024,4,false>(in,out,bytes,16); run<528,4,false>(in,out,bytes,16);
    return 0;
}
