// PNG input/output for Image<Color<float>>.  API of the reference's include/PathTrace/image/image_io.h; implemented
// on zlib directly (cpupathtrace_b200/host/image_io.cpp) because libpng is not part of this toolchain.
#ifndef PATHTRACE_IMAGE_IO_H
#define PATHTRACE_IMAGE_IO_H

#include <PathTrace/image/image.h>
#include <PathTrace/util/color.h>

#include <filesystem>
#include <istream>
#include <ostream>
#include <string>

namespace io {

    //! Decodes an 8-bit RGB / RGBA (or grey / grey-alpha / palette) PNG; channels are mapped to [0, 1].
    //! @throw std::logic_error when decoding fails
    Image<Color<float>> readRGBImage(std::basic_istream<char> &stream) noexcept(false);
    Image<Color<float>> readRGBImage(const std::string &path) noexcept(false);
    Image<Color<float>> readRGBImage(const std::filesystem::path &path) noexcept(false);

    //! Encodes as non-interlaced RGBA8; channels are mapped from [0, 1] by round(255 v) and clamped.
    //! @throw std::logic_error when encoding fails
    void writeRGBImage(std::basic_ostream<char> &stream, const Image<Color<float>> &image) noexcept(false);
    void writeRGBImage(const std::string &path, const Image<Color<float>> &image) noexcept(false);
    void writeRGBImage(const std::filesystem::path &path, const Image<Color<float>> &image) noexcept(false);

}

#endif /* PATHTRACE_IMAGE_IO_H */
