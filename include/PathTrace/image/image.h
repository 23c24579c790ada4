// Row-major 2D grid of values; the output type of processItem / processJob.
// API of the reference's include/PathTrace/image/image.h.
#ifndef PATHTRACE_IMAGE_H
#define PATHTRACE_IMAGE_H

#include <PathTrace/util/color.h>

#include <cassert>
#include <cstddef>
#include <vector>

template<typename T = Color<float>>
class Image {
  public:
    using value_type = T;

    Image() = default;

    //! width x height cells, value-initialised
    Image(int width, int height) : width(width), height(height), data_(static_cast<std::size_t>(width) * static_cast<std::size_t>(height)) {}

    T operator()(int x, int y) const noexcept { return data_[index(x, y)]; }
    T &operator()(int x, int y) noexcept { return data_[index(x, y)]; }

    std::size_t size() const noexcept { return data_.size(); }

    const T *data() const noexcept { return data_.data(); }
    T *data() noexcept { return data_.data(); }

    int getWidth() const noexcept { return width; }
    int getHeight() const noexcept { return height; }

  protected:
    void assertContainsPoint([[maybe_unused]] int x, [[maybe_unused]] int y) const noexcept { assert(x >= 0 && x < width && y >= 0 && y < height); }

  private:
    std::size_t index(int x, int y) const noexcept {
        assertContainsPoint(x, y);
        return static_cast<std::size_t>(y) * static_cast<std::size_t>(width) + static_cast<std::size_t>(x);
    }

    int width = 0;
    int height = 0;
    std::vector<T> data_;
};

#endif /* PATHTRACE_IMAGE_H */
