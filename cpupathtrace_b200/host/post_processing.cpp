// Tone mapping and gamma correction behind the reference's API (include/PathTrace/post_processing.h).
// The arithmetic runs on the device (cpupathtrace_b200/csrc/post_process.cuh via ptb_post_process) and is bit-exact
// with the reference's src/post_processing.cpp; a frame that is already in HBM can be processed there directly through
// the C-ABI with PTB_FLAG_DEVICE_IO.
#include "device.h"

#include <PathTrace/post_processing.h>

namespace {

    void run(Image<> &image, uint32_t mode, float gamma) {
        if(image.size() == 0) {
            return;
        }
        static_assert(sizeof(Color<float>) == 4 * sizeof(float), "Color<float> must be four packed floats");
        ptb::host::check(ptb_post_process(ptb::host::defaultContext(), reinterpret_cast<float *>(image.data()), image.getWidth(), image.getHeight(), mode, gamma,
                                          0U),
                         "post-processing");
    }

}

void toneMap(Image<> &image) {
    run(image, PTB_POST_TONE_MAP, 1.0F);
}

void gammaCorrect(Image<> &image, float gamma) {
    run(image, PTB_POST_GAMMA, gamma);
}

void postProcess(Image<> &image) {
    // toneMap(image); gammaCorrect(image) with the default gamma 1.8 (post_processing.cpp:179-182)
    run(image, PTB_POST_BOTH, 1.8F);
}
