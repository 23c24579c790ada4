// Tone mapping and gamma correction (host side).  Follows the algorithm of the reference's src/post_processing.cpp:
// a histogram-equalising remap of a brightness heuristic through 1024 gaussian-weighted segments, then a
// hue-preserving gamma pre-correction.  Operates on RGB only; alpha is left untouched.
#include <PathTrace/post_processing.h>

#include <PathTrace/base.h>

#include <algorithm>
#include <cmath>
#include <limits>
#include <vector>

namespace {

    float peakChannel(const Color<float> &c) {
        return std::max({c[0], c[1], c[2]});
    }

    // alpha * (mean(rgb) + max(rgb)) / 2
    float brightnessHeuristic(const Color<float> &c) {
        return c[3] * ((c[0] + c[1] + c[2]) / 3.0F + std::max({c[0], c[1], c[2]})) / 2.0F;
    }

    float normalDensity(float t, float mu, float sigma) {
        constexpr float pi = static_cast<float>(M_PI);
        const float scale = 1.0F / (std::sqrt(2 * pi));
        const float z = (t - mu) / (sigma);
        return scale * std::exp(-(z * z) / 2.0F) / sigma;
    }

}

void toneMap(Image<> &image) {
    const int pixel_count = image.getWidth() * image.getHeight();
    if(pixel_count <= 0) {
        return;
    }
    Color<float> *pixels = image.data();

    // 1. range of the brightness heuristic (the range always includes 0 and at least 1e-4)
    float lowest = 0.0F;
    float highest = 1E-4F;
    std::vector<float> sorted(static_cast<std::size_t>(pixel_count));
    for(int i = 0; i < pixel_count; i++) {
        const float b = brightnessHeuristic(pixels[i]);
        sorted[i] = b;
        lowest = std::min(lowest, b);
        highest = std::max(highest, b);
    }

    // 2. all brightness values in ascending order (the reference buckets into 1024 bins, sorts each and concatenates,
    //    which is a full sort)
    std::sort(sorted.begin(), sorted.end());

    // 3. gaussian-weighted share of the pixels for each output segment
    const int segments = std::min(1024, pixel_count);
    std::vector<float> weight(static_cast<std::size_t>(segments));
    float total_weight = 0.0F;
    for(int s = 0; s < segments; s++) {
        float centre = (static_cast<float>(s) + 0.5F) / static_cast<float>(segments);
        centre = 2.0F * (centre - 0.5F);
        weight[s] = 0.1F + normalDensity(centre, 0.0F, 0.3F);
        total_weight += weight[s];
    }

    // 4. upper brightness bound ("ceiling") of every segment, carrying rounding remainders forward
    std::vector<float> ceiling;
    ceiling.reserve(static_cast<std::size_t>(segments));
    int consumed = 0;
    float carried = 0.0F;
    for(int s = 0; s < segments - 1; s++) {
        const int share = static_cast<int>(std::round(weight[s] * static_cast<float>(pixel_count) / total_weight + carried));
        if(share > 0) {
            const int last = std::min(consumed + share - 1, pixel_count - 1);
            ceiling.push_back(sorted[static_cast<std::size_t>(last)]);
            consumed += share;
            carried = 0.0F;
        }
        else {
            ceiling.push_back(s > 0 ? ceiling[static_cast<std::size_t>(s) - 1] : lowest);
            carried += weight[s] * static_cast<float>(pixel_count) / total_weight;
        }
    }
    ceiling.push_back(highest);

    // 5. remap every pixel: position inside its input segment -> same position inside the equal-width output segment
    constexpr float tiny = std::numeric_limits<float>::min();
    for(int i = 0; i < pixel_count; i++) {
        Color<float> &pixel = pixels[i];
        const float peak = std::max(peakChannel(pixel), tiny);
        const float b = brightnessHeuristic(pixel);

        const auto found = std::lower_bound(ceiling.begin(), ceiling.end(), b);
        const int s = found == ceiling.end() ? segments - 1 : static_cast<int>(found - ceiling.begin());
        const float upper = ceiling[static_cast<std::size_t>(s)];
        const float lower = s > 0 ? ceiling[static_cast<std::size_t>(s) - 1] : lowest;
        const float span = std::max(upper - lower, tiny);
        const float position = (b - lower) / span;

        const float out_upper = static_cast<float>(s + 1) / static_cast<float>(segments);
        const float out_lower = static_cast<float>(s) / static_cast<float>(segments);
        const float mapped = out_lower + position * (out_upper - out_lower);

        const float factor = mapped / peak;
        pixel[0] *= factor;
        pixel[1] *= factor;
        pixel[2] *= factor;
    }
}

void gammaCorrect(Image<> &image, float gamma) {
    const int pixel_count = image.getWidth() * image.getHeight();
    Color<float> *pixels = image.data();
    for(int i = 0; i < pixel_count; i++) {
        const float factor = std::pow(peakChannel(pixels[i]), 1.0F / gamma - 1.0F);
        pixels[i][0] *= factor;
        pixels[i][1] *= factor;
        pixels[i][2] *= factor;
    }
}

void postProcess(Image<> &image) {
    toneMap(image);
    gammaCorrect(image);
}
