/*
 * amx_color.cu -- K7: colour-space ingest.  One pass over a key frame that produces, per canvas
 * position, the colour as the reference STORES it (RGB->HSP when blob_delimiter == HSP,
 * morph.cpp:299-301) and the colour as the reference FETCHES it (HSP->RGB, morph.cpp:387-389):
 * an 8-bit lossy round trip that the renderer must reproduce.  Absent positions become 0.
 * HBM-bound: 4 B + 1 B read, 8 B written per position.
 */
#include "amx_engine.h"

namespace amx {

__global__ void __launch_bounds__(256)
k_color_ingest(const uint32_t *__restrict__ raw, const uint8_t *__restrict__ present, uint32_t *__restrict__ stored,
               uint32_t *__restrict__ fetch, size_t n, int hsp, unsigned long long *__restrict__ count) {
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    unsigned has = 0;
    if (i < n) {
        has = present[i] ? 1u : 0u;
        uint32_t s = 0, f = 0;
        if (has) {
            uint32_t c = raw[i];
            s = hsp ? rgb_to_hsp(c) : c;
            f = hsp ? hsp_to_rgb(s) : c;
        }
        stored[i] = s;
        fetch[i] = f;
    }
    unsigned ballot = __ballot_sync(0xffffffffu, has);
    if ((threadIdx.x & 31) == 0 && ballot) atomicAdd(count, (unsigned long long) __popc(ballot));
}

int engine_upload_convert(Engine *E, uint32_t index, const uint32_t *d_raw) {
    FrameDev &f = E->frames[index];
    size_t n = E->canvas();
    unsigned long long *d_count = nullptr;
    if (!dev_alloc(E, (void **) &d_count, 8, "count")) return AMX_ERR_NOMEM;
    cudaMemsetAsync(d_count, 0, 8, E->stream);
    k_color_ingest<<<div_up(n, 256), 256, 0, E->stream>>>(d_raw, f.present, f.stored, f.fetch, n, E->p.blob_delimiter == K_HSP, d_count);
    E->launches++;
    unsigned long long cnt = 0;
    cudaMemcpyAsync(&cnt, d_count, 8, cudaMemcpyDeviceToHost, E->stream);
    cudaError_t e = cudaStreamSynchronize(E->stream);
    dev_free(d_count);
    if (E->fail(e, "color ingest") || E->check("color ingest")) return AMX_ERR_CUDA;
    f.pixel_count = cnt;
    return AMX_OK;
}

} // namespace amx
