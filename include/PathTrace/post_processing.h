// Tone mapping and gamma correction of a rendered image (host side; API of the reference's
// include/PathTrace/post_processing.h).
#ifndef PATHTRACE_POST_PROCESSING_H
#define PATHTRACE_POST_PROCESSING_H

#include <PathTrace/image/image.h>

//! Histogram-equalising map of an arbitrary finite value range to [0, 1], in place (RGB only)
void toneMap(Image<> &image);

//! Pre-corrects for a display gamma, preserving hue: rgb *= max(rgb)^(1/gamma - 1), in place
void gammaCorrect(Image<> &image, float gamma = 1.8F);

//! toneMap followed by gammaCorrect
void postProcess(Image<> &image);

#endif // PATHTRACE_POST_PROCESSING_H
