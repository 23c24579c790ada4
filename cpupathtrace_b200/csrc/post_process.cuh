// Tone mapping and gamma correction on the device (reference src/post_processing.cpp:32-182), so that an image
// rendered into HBM can be post-processed without a host round trip.  Bit-exact with the reference: every per-pixel
// operation is fp32 +,-,*,/ (IEEE, no contraction), the global sort is order-exact, the gamma factor uses the glibc powf
// restatement, and the 1024 segment weights (std::exp on the host in the reference) are computed on the host here too.
#ifndef PTB_POST_PROCESS_CUH
#define PTB_POST_PROCESS_CUH

#include "glibc_libm.cuh"

namespace ptb {

    // getBrightness (post_processing.cpp:22-25): largest colour channel
    PTB_DEV float peakChannel(float4 c) {
        return stdmax(stdmax(c.x, c.y), c.z);
    }

    // getBrightnessHeuristic (post_processing.cpp:27-30): alpha * (mean(rgb) + max(rgb)) / 2
    PTB_DEV float brightnessHeuristic(float4 c) {
        return c.w * ((c.x + c.y + c.z) / 3.0F + stdmax(stdmax(c.x, c.y), c.z)) / 2.0F;
    }

    __global__ void brightnessKernel(const float4 *__restrict__ pixels, uint32_t n, float *__restrict__ brightness) {
        const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
        if(i < n) {
            brightness[i] = brightnessHeuristic(pixels[i]);
        }
    }

    // range[0] = min(0, all), range[1] = max(1e-4, all)  (post_processing.cpp:35-47); in: device min / max of the values
    __global__ void rangeKernel(float *__restrict__ range) {
        if(blockIdx.x == 0U && threadIdx.x == 0U) {
            range[0] = stdmin(0.0F, range[0]);
            range[1] = stdmax(1E-4F, range[1]);
        }
    }

    // Segment ceilings (post_processing.cpp:107-128).  pick[i] >= 0: index into the sorted brightness values;
    // pick[i] < 0: repeat the previous ceiling (the first one falls back to the minimum brightness).
    __global__ void ceilingsKernel(const float *__restrict__ sorted, const int32_t *__restrict__ pick, int32_t segments, const float *__restrict__ range,
                                   float *__restrict__ ceilings) {
        if(blockIdx.x != 0U || threadIdx.x != 0U) {
            return;
        }
        for(int32_t i = 0; i < segments - 1; i++) {
            if(pick[i] >= 0) {
                ceilings[i] = sorted[pick[i]];
            }
            else {
                ceilings[i] = i > 0 ? ceilings[i - 1] : range[0];
            }
        }
        ceilings[segments - 1] = range[1];
    }

    // per-pixel remap (post_processing.cpp:130-162)
    __global__ void toneMapKernel(float4 *__restrict__ pixels, uint32_t n, const float *__restrict__ ceilings, int32_t segments,
                                  const float *__restrict__ range) {
        const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
        if(i >= n) {
            return;
        }
        float4 pixel = pixels[i];
        const float tiny = 1.17549435e-38F; // std::numeric_limits<float>::min()
        const float brightness = stdmax(peakChannel(pixel), tiny);
        const float heuristic = brightnessHeuristic(pixel);

        // std::lower_bound(ceilings, heuristic)
        int32_t lo = 0;
        int32_t len = segments;
        while(len > 0) {
            const int32_t half = len >> 1;
            if(ceilings[lo + half] < heuristic) {
                lo += half + 1;
                len -= half + 1;
            }
            else {
                len = half;
            }
        }
        const int32_t index = lo < segments ? lo : segments - 1;
        const float upper = ceilings[index];
        const float lower = index > 0 ? ceilings[index - 1] : range[0];
        const float span = stdmax(upper - lower, tiny);
        const float value = (heuristic - lower) / span;
        const float mapped_upper = static_cast<float>(index + 1) / static_cast<float>(segments);
        const float mapped_lower = static_cast<float>(index) / static_cast<float>(segments);
        const float mapped_span = mapped_upper - mapped_lower;
        const float mapped_value = mapped_lower + value * mapped_span;
        const float factor = mapped_value / brightness;
        pixel.x *= factor;
        pixel.y *= factor;
        pixel.z *= factor;
        pixels[i] = pixel;
    }

    // gammaCorrect (post_processing.cpp:165-177): rgb *= max(rgb) ^ (1 / gamma - 1); exponent computed by the host
    __global__ void gammaKernel(float4 *__restrict__ pixels, uint32_t n, float exponent) {
        const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
        if(i >= n) {
            return;
        }
        float4 pixel = pixels[i];
        const float brightness = peakChannel(pixel);
        float factor;
        if(exponent == 0.0F) {
            factor = (brightness != brightness) ? brightness : 1.0F; // powf(x, 0) == 1
        }
        else {
            factor = glibcPowfPositive(brightness, exponent);
        }
        pixel.x *= factor;
        pixel.y *= factor;
        pixel.z *= factor;
        pixels[i] = pixel;
    }

}

#endif
