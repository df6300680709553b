/*
oracle/stub_gil/boost/gil.hpp -- TEST INFRASTRUCTURE, not product code.

A stand-in for the slice of Boost.GIL that the reference's renderers/image_renderer.hpp and
utils/image.hpp touch (Boost is not installed in this image): the five image typedefs, their
(width, height) constructor, view(img).row_begin(y) and pixel assignment. With it the
UNMODIFIED image_renderer.hpp compiles (oracle/Makefile, target refimg) so that the GPU tone
map can be checked against the reference's own pixel arithmetic. PNG encoding is not part of
this stub; the harness captures the pixels instead of writing a file.
*/
#pragma once

#include <cstddef>
#include <cstdint>
#include <type_traits>
#include <vector>

namespace boost { namespace gil {

template <typename P, int C>
struct pixel
{
    P v[C];
    pixel() { for (int i = 0; i < C; ++i) v[i] = 0; }
    pixel(P a, P b, P c) { static_assert(C == 3,"rgb only"); v[0] = a; v[1] = b; v[2] = c; }
    template <typename A, typename = typename std::enable_if<std::is_arithmetic<A>::value && C == 1>::type>
    pixel& operator=(A a) { v[0] = (P)a; return *this; }
};

template <typename Pix>
struct image_view
{
    Pix *base;
    std::size_t w, h;
    Pix *row_begin(std::size_t y) const { return base + y*w; }
};

template <typename P, int C>
struct image
{
    typedef pixel<P,C> pixel_t;
    typedef P channel_t;
    static constexpr int channels = C;
    std::size_t w, h;
    std::vector<pixel_t> data;
    image(): w(0), h(0) {}
    image(std::size_t w_, std::size_t h_): w(w_), h(h_), data(w_*h_) {}
    std::size_t width() const { return w; }
    std::size_t height() const { return h; }
};

template <typename P, int C>
image_view<pixel<P,C>> view(image<P,C>& img)
{
    return image_view<pixel<P,C>>{img.data.data(),img.w,img.h};
}

typedef image<std::uint8_t,1> gray1_image_t;   /* one bit per pixel in GIL; only named here */
typedef image<std::uint8_t,1> gray8_image_t;
typedef image<std::uint16_t,1> gray16_image_t;
typedef image<std::uint8_t,3> rgb8_image_t;
typedef image<std::uint16_t,3> rgb16_image_t;

}} // namespace boost::gil
