#include "jpeg.hpp"

#include <cmath>
#include <cstdio>
#include <cstring>

namespace cthost {
namespace {

const uint8_t kZigzag[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                             35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
// ITU-T T.81 Annex K.1 / K.2, natural (row-major) order
const int kLumQ[64] = {16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
                       18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
const int kChrQ[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
                       99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
// Annex K.3 Huffman tables
const uint8_t kDcLumBits[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
const uint8_t kDcChrBits[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
const uint8_t kDcVals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
const uint8_t kAcLumBits[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
const uint8_t kAcLumVals[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1, 0x08,
    0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28,
    0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59,
    0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89,
    0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6,
    0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2,
    0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};
const uint8_t kAcChrBits[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
const uint8_t kAcChrVals[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42, 0x91,
    0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26,
    0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58,
    0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87,
    0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4,
    0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda,
    0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};

struct Huff { uint16_t code[256]; uint8_t size[256]; };

void build_huff(const uint8_t bits[16], const uint8_t *vals, Huff &h) {
  memset(&h, 0, sizeof h);
  unsigned code = 0;
  int k = 0;
  for (int len = 1; len <= 16; len++) {
    for (int i = 0; i < bits[len - 1]; i++) { h.code[vals[k]] = (uint16_t)code++; h.size[vals[k]] = (uint8_t)len; k++; }
    code <<= 1;
  }
}

struct BitWriter {
  std::vector<uint8_t> &out;
  uint32_t acc = 0;
  int n = 0;
  explicit BitWriter(std::vector<uint8_t> &o) : out(o) {}
  void put(unsigned code, int size) {
    acc = (acc << size) | (code & ((1u << size) - 1u));
    n += size;
    while (n >= 8) {
      uint8_t b = (uint8_t)(acc >> (n - 8));
      out.push_back(b);
      if (b == 0xFF) out.push_back(0);
      n -= 8;
    }
  }
  void flush() { if (n) put(0x7F, 8 - n); }
};

void fdct8x8(const float in[64], float out[64], const float cs[64]) {
  float tmp[64];
  for (int y = 0; y < 8; y++)
    for (int u = 0; u < 8; u++) {
      float s = 0;
      for (int x = 0; x < 8; x++) s += in[y * 8 + x] * cs[u * 8 + x];
      tmp[y * 8 + u] = s;
    }
  for (int u = 0; u < 8; u++)
    for (int v = 0; v < 8; v++) {
      float s = 0;
      for (int y = 0; y < 8; y++) s += tmp[y * 8 + u] * cs[v * 8 + y];
      out[v * 8 + u] = s;
    }
}

void put16(std::vector<uint8_t> &o, int v) { o.push_back((uint8_t)(v >> 8)); o.push_back((uint8_t)v); }

int encode_block(BitWriter &bw, const float blk[64], const int q[64], const float cs[64], int prev_dc, const Huff &dc, const Huff &ac) {
  float coef[64];
  fdct8x8(blk, coef, cs);
  int zz[64];
  for (int i = 0; i < 64; i++) {
    float v = coef[kZigzag[i]] / (float)q[kZigzag[i]];
    zz[i] = (int)(v < 0 ? v - 0.5f : v + 0.5f);
  }
  auto category = [](int v) { int a = v < 0 ? -v : v, c = 0; while (a) { c++; a >>= 1; } return c; };
  auto bits_of = [](int v, int c) { return (unsigned)(v < 0 ? v + (1 << c) - 1 : v); };
  int diff = zz[0] - prev_dc;
  int c = category(diff);
  bw.put(dc.code[c], dc.size[c]);
  if (c) bw.put(bits_of(diff, c), c);
  int last = 63;
  while (last > 0 && zz[last] == 0) last--;
  int run = 0;
  for (int i = 1; i <= last; i++) {
    if (zz[i] == 0) { run++; continue; }
    while (run >= 16) { bw.put(ac.code[0xF0], ac.size[0xF0]); run -= 16; }
    int cc = category(zz[i]);
    int sym = (run << 4) | cc;
    bw.put(ac.code[sym], ac.size[sym]);
    bw.put(bits_of(zz[i], cc), cc);
    run = 0;
  }
  if (last != 63) bw.put(ac.code[0x00], ac.size[0x00]);
  return zz[0];
}

}  // namespace

void encode_jpeg(std::vector<uint8_t> &o, int w, int h, const uint8_t *rgb, int quality) {
  o.clear();
  if (quality < 1) quality = 1;
  if (quality > 100) quality = 100;
  int scale = quality < 50 ? 5000 / quality : 200 - quality * 2;
  int ql[64], qc[64];
  for (int i = 0; i < 64; i++) {
    int a = (kLumQ[i] * scale + 50) / 100, b = (kChrQ[i] * scale + 50) / 100;
    ql[i] = a < 1 ? 1 : (a > 255 ? 255 : a);
    qc[i] = b < 1 ? 1 : (b > 255 ? 255 : b);
  }
  float cs[64];  // cs[u*8+x] = c(u)/2 * cos((2x+1) u pi / 16)
  for (int u = 0; u < 8; u++)
    for (int x = 0; x < 8; x++) cs[u * 8 + x] = (float)((u == 0 ? std::sqrt(0.125) : 0.5) * std::cos((2 * x + 1) * u * M_PI / 16.0));
  Huff dcl, dcc, acl, acc;
  build_huff(kDcLumBits, kDcVals, dcl); build_huff(kDcChrBits, kDcVals, dcc);
  build_huff(kAcLumBits, kAcLumVals, acl); build_huff(kAcChrBits, kAcChrVals, acc);

  // headers
  o.push_back(0xFF); o.push_back(0xD8);
  const uint8_t app0[] = {0xFF, 0xE0, 0, 16, 'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0};
  o.insert(o.end(), app0, app0 + sizeof app0);
  o.push_back(0xFF); o.push_back(0xDB); put16(o, 2 + 2 * 65);
  o.push_back(0); for (int i = 0; i < 64; i++) o.push_back((uint8_t)ql[kZigzag[i]]);
  o.push_back(1); for (int i = 0; i < 64; i++) o.push_back((uint8_t)qc[kZigzag[i]]);
  o.push_back(0xFF); o.push_back(0xC0); put16(o, 17); o.push_back(8); put16(o, h); put16(o, w); o.push_back(3);
  o.push_back(1); o.push_back(0x22); o.push_back(0);
  o.push_back(2); o.push_back(0x11); o.push_back(1);
  o.push_back(3); o.push_back(0x11); o.push_back(1);
  o.push_back(0xFF); o.push_back(0xC4); put16(o, 2 + (17 + 12) * 2 + (17 + 162) * 2);
  o.push_back(0x00); o.insert(o.end(), kDcLumBits, kDcLumBits + 16); o.insert(o.end(), kDcVals, kDcVals + 12);
  o.push_back(0x10); o.insert(o.end(), kAcLumBits, kAcLumBits + 16); o.insert(o.end(), kAcLumVals, kAcLumVals + 162);
  o.push_back(0x01); o.insert(o.end(), kDcChrBits, kDcChrBits + 16); o.insert(o.end(), kDcVals, kDcVals + 12);
  o.push_back(0x11); o.insert(o.end(), kAcChrBits, kAcChrBits + 16); o.insert(o.end(), kAcChrVals, kAcChrVals + 162);
  const uint8_t sos[] = {0xFF, 0xDA, 0, 12, 3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0};
  o.insert(o.end(), sos, sos + sizeof sos);

  BitWriter bw(o);
  int dcy = 0, dcb = 0, dcr = 0;
  for (int my = 0; my < h; my += 16) {
    for (int mx = 0; mx < w; mx += 16) {
      float Y[256], Cb[256], Cr[256];
      for (int y = 0; y < 16; y++) {
        int sy = my + y < h ? my + y : h - 1;
        for (int x = 0; x < 16; x++) {
          int sx = mx + x < w ? mx + x : w - 1;
          const uint8_t *p = rgb + 3 * ((size_t)sy * w + sx);
          float r = p[0], g = p[1], b = p[2];
          Y[y * 16 + x] = 0.299f * r + 0.587f * g + 0.114f * b - 128.0f;
          Cb[y * 16 + x] = -0.168736f * r - 0.331264f * g + 0.5f * b;
          Cr[y * 16 + x] = 0.5f * r - 0.418688f * g - 0.081312f * b;
        }
      }
      float blk[64];
      for (int by = 0; by < 2; by++)
        for (int bx = 0; bx < 2; bx++) {
          for (int y = 0; y < 8; y++)
            for (int x = 0; x < 8; x++) blk[y * 8 + x] = Y[(by * 8 + y) * 16 + bx * 8 + x];
          dcy = encode_block(bw, blk, ql, cs, dcy, dcl, acl);
        }
      for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++) {
          int i = (2 * y) * 16 + 2 * x;
          blk[y * 8 + x] = 0.25f * (Cb[i] + Cb[i + 1] + Cb[i + 16] + Cb[i + 17]);
        }
      dcb = encode_block(bw, blk, qc, cs, dcb, dcc, acc);
      for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++) {
          int i = (2 * y) * 16 + 2 * x;
          blk[y * 8 + x] = 0.25f * (Cr[i] + Cr[i + 1] + Cr[i + 16] + Cr[i + 17]);
        }
      dcr = encode_block(bw, blk, qc, cs, dcr, dcc, acc);
    }
  }
  bw.flush();
  o.push_back(0xFF); o.push_back(0xD9);
}

bool write_jpeg(const std::string &path, int w, int h, const uint8_t *rgb, int quality) {
  if (w <= 0 || h <= 0 || w > 65535 || h > 65535) return false;   // baseline JPEG stores the dimensions in 16 bits (SOF0)
  std::vector<uint8_t> buf;
  encode_jpeg(buf, w, h, rgb, quality);
  FILE *f = fopen(path.c_str(), "wb");
  if (!f) return false;
  bool ok = fwrite(buf.data(), 1, buf.size(), f) == buf.size();
  return fclose(f) == 0 && ok;
}

}  // namespace cthost
