// color_flow -- command-line twin of the Middlebury tool the reference vendors (middlebury/flow-code/color_flow.cpp:68-98):
//
//     color_flow [-quiet] in.flo out.png [maxmotion]
//
// reads a .flo file (Flow::ReadFlowFile), colour-codes it (Flow::MotionToColor, rw_flow.cpp:202-300 -- the class the
// reference's main() uses for flow.png, main_class.cpp:73-75) and writes a PNG (or a binary PPM when the name ends in .ppm).
// Same console output as the original: the motion-range line on stdout, "normalizing by" on stderr unless -quiet.
// Host-only: links libbbme.so for the codec and the colour wheel; needs no GPU.  The PNG is written with stored (uncompressed)
// deflate blocks, so no libpng / zlib is needed.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "bbme.h"

static const char* usage = "\n  usage: %s [-quiet] in.flo out.png [maxmotion]\n";

static uint32_t crc_table[256];
static void crc_init() {
  for (uint32_t n = 0; n < 256; ++n) {
    uint32_t c = n;
    for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xedb88320u ^ (c >> 1) : c >> 1;
    crc_table[n] = c;
  }
}
static uint32_t crc_update(uint32_t c, const uint8_t* p, size_t n) {
  for (size_t i = 0; i < n; ++i) c = crc_table[(c ^ p[i]) & 0xff] ^ (c >> 8);
  return c;
}
static void put32(std::vector<uint8_t>& v, uint32_t x) {
  v.push_back((uint8_t)(x >> 24)); v.push_back((uint8_t)(x >> 16)); v.push_back((uint8_t)(x >> 8)); v.push_back((uint8_t)x);
}
static void chunk(FILE* f, const char* type, const std::vector<uint8_t>& data) {
  std::vector<uint8_t> head;
  put32(head, (uint32_t)data.size());
  fwrite(head.data(), 1, 4, f);
  fwrite(type, 1, 4, f);
  if (!data.empty()) fwrite(data.data(), 1, data.size(), f);
  uint32_t c = crc_update(0xffffffffu, reinterpret_cast<const uint8_t*>(type), 4);
  c = crc_update(c, data.data(), data.size()) ^ 0xffffffffu;
  std::vector<uint8_t> tail;
  put32(tail, c);
  fwrite(tail.data(), 1, 4, f);
}

// 8-bit RGB PNG from BGR rows (OpenCV's channel order, what Flow::MotionToColor produces)
static bool write_png(const char* path, const uint8_t* bgr, int w, int h) {
  FILE* f = fopen(path, "wb");
  if (!f) return false;
  crc_init();
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  fwrite(sig, 1, 8, f);
  std::vector<uint8_t> ihdr;
  put32(ihdr, (uint32_t)w);
  put32(ihdr, (uint32_t)h);
  ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
  chunk(f, "IHDR", ihdr);
  std::vector<uint8_t> raw;  // filter byte 0 + RGB per row
  raw.reserve((size_t)h * (3 * (size_t)w + 1));
  for (int y = 0; y < h; ++y) {
    raw.push_back(0);
    for (int x = 0; x < w; ++x) {
      const uint8_t* p = bgr + ((size_t)y * w + x) * 3;
      raw.push_back(p[2]); raw.push_back(p[1]); raw.push_back(p[0]);
    }
  }
  std::vector<uint8_t> z;
  z.push_back(0x78); z.push_back(0x01);
  uint32_t a = 1, b = 0;
  for (size_t off = 0; off < raw.size() || off == 0;) {
    const size_t n = raw.size() - off < 65535 ? raw.size() - off : 65535;
    const bool last = off + n >= raw.size();
    z.push_back(last ? 1 : 0);
    z.push_back((uint8_t)(n & 0xff)); z.push_back((uint8_t)(n >> 8));
    z.push_back((uint8_t)(~n & 0xff)); z.push_back((uint8_t)((~n >> 8) & 0xff));
    for (size_t i = 0; i < n; ++i) {
      z.push_back(raw[off + i]);
      a = (a + raw[off + i]) % 65521u;
      b = (b + a) % 65521u;
    }
    off += n;
    if (last) break;
  }
  put32(z, (b << 16) | a);
  chunk(f, "IDAT", z);
  chunk(f, "IEND", std::vector<uint8_t>());
  return fclose(f) == 0;
}

static bool write_ppm(const char* path, const uint8_t* bgr, int w, int h) {
  FILE* f = fopen(path, "wb");
  if (!f) return false;
  fprintf(f, "P6\n%d %d\n255\n", w, h);
  for (size_t i = 0; i < (size_t)w * h; ++i) {
    const uint8_t rgb[3] = {bgr[3 * i + 2], bgr[3 * i + 1], bgr[3 * i]};
    fwrite(rgb, 1, 3, f);
  }
  return fclose(f) == 0;
}

int main(int argc, char* argv[]) {
  int verbose = 1, argn = 1;
  if (argc > 1 && argv[1][0] == '-' && argv[1][1] == 'q') {
    verbose = 0;
    argn++;
  }
  if (!(argn >= argc - 3 && argn <= argc - 2)) {
    fprintf(stderr, usage, argv[0]);
    fprintf(stderr, "\n");
    return -1;
  }
  const char* flowname = argv[argn++];
  const char* outname = argv[argn++];
  const float maxmotion = argn < argc ? (float)atof(argv[argn++]) : -1.f;
  int w = 0, h = 0;
  int rc = bbme_flo_read_header(flowname, &w, &h);
  if (rc != BBME_OK) {
    fprintf(stderr, "ReadFlowFile(%s): %s\n", flowname, bbme_status_string(rc));
    return -1;
  }
  std::vector<float> flow((size_t)w * h * 2);
  if ((rc = bbme_flo_read(flowname, flow.data(), w, h)) != BBME_OK) {
    fprintf(stderr, "ReadFlowFile(%s): %s\n", flowname, bbme_status_string(rc));
    return -1;
  }
  std::vector<uint8_t> bgr((size_t)w * h * 3);
  float r[5];
  if (bbme_flow_to_color(flow.data(), w, h, maxmotion, bgr.data(), r) != BBME_OK) return -1;
  printf("max motion: %.4f  motion range: u = %.3f .. %.3f;  v = %.3f .. %.3f\n", r[0], r[1], r[2], r[3], r[4]);
  float maxrad = maxmotion > 0 ? maxmotion : r[0];
  if (maxrad == 0) maxrad = 1;
  if (verbose) fprintf(stderr, "normalizing by %g\n", maxrad);
  const size_t len = strlen(outname);
  const bool ppm = len > 4 && strcmp(outname + len - 4, ".ppm") == 0;
  if (!(ppm ? write_ppm(outname, bgr.data(), w, h) : write_png(outname, bgr.data(), w, h))) {
    fprintf(stderr, "cannot write %s\n", outname);
    return -1;
  }
  if (verbose) fprintf(stderr, "Writing image %s\n", outname);
  return 0;
}
