// jpeg.hpp — baseline JPEG writer for the three output images.
// The reference hands its 8-bit RGB buffers to stb_image_write's stbi_write_jpg(..., 3, data, 90)
// (inc/images.hpp:39,64,86); stb is not vendored and there is no libjpeg offline, so this is an own
// baseline encoder with the same parameter choices stb makes at quality 90: JFIF, YCbCr, 4:2:0 chroma
// subsampling (stb subsamples for quality <= 90), Annex-K quantisation tables scaled by (200 - 2q)/100,
// Annex-K Huffman tables.  Byte streams are not expected to be identical to stb's; decoded images are.
#ifndef CUTRACE_B200_HOST_JPEG_HPP
#define CUTRACE_B200_HOST_JPEG_HPP
#include <cstdint>
#include <string>
#include <vector>

namespace cthost {
// rgb: w*h*3 bytes, row-major, row 0 = top. Returns false if the file cannot be written.
bool write_jpeg(const std::string &path, int w, int h, const uint8_t *rgb, int quality = 90);
void encode_jpeg(std::vector<uint8_t> &out, int w, int h, const uint8_t *rgb, int quality = 90);
}  // namespace cthost
#endif
