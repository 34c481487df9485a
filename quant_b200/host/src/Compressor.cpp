// Codec layer - replaces /root/reference/src/Compressor.cpp.
//
// compress() is the drop-in cut: where the reference expands the image into N x dim
// doubles (getBlocksAsVectorsFromImage, :31-62) and calls quantize() (:116-122), this one passes the
// image's own byte buffer to libqb200, whose kernels gather block vectors with the same layout rule.
// Block extraction / decode / .quant I/O below are host restatements kept for API completeness and
// for reading and writing files; none of them is on the accelerated path.
#include "Compressor.hpp"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <exception>
#include <fstream>
#include <iomanip>
#include <sstream>
#include <stdexcept>
#include <thread>

#include "../../../include/qb200.h"
#include "B200Context.hpp"

namespace {

size_t ceil_div(size_t a, size_t b) { return (a + b - 1) / b; }

// floor(log2(n)) for n >= 1 (src/Compressor.cpp:167-172): bits per stored index.
size_t index_bits(size_t n) {
  size_t bits = 0;
  while (n >>= 1) bits++;
  return bits;
}

// Visits every (vector, element, pixel) triple of the block layout: vector i*hBlocks + j holds, at
// element ((x - i*w)*h + (y - j*h))*3 + channel, pixel x*ySize + y.  The image buffer is addressed as
// [xSize][ySize] although a PPM is row-major in xSize - the reference does the same, and blocks that
// overflow in y wrap into the next line instead of being padded.
template <class F>
void for_each_block_pixel(size_t xSize, size_t ySize, size_t w, size_t h, F f) {
  const size_t wB = ceil_div(xSize, w), hB = ceil_div(ySize, h);
  for (size_t i = 0; i < wB; i++)
    for (size_t j = 0; j < hB; j++)
      for (size_t dx = 0; dx < w; dx++)
        for (size_t dy = 0; dy < h; dy++)
          f(i * hB + j, (dx * h + dy) * 3, (i * w + dx) * ySize + (j * h + dy));
}

std::string pretty_bytes(size_t bytes) {
  std::ostringstream s;
  if (bytes < 1024)
    s << bytes << "b";
  else if (bytes < 1024 * 1024)
    s << bytes / 1024 << "," << bytes % 1024 << "Kb";
  else  // the remainder is printed in bytes, as the reference does
    s << bytes / (1024 * 1024) << "," << bytes % (1024 * 1024) << "Mb";
  return s.str();
}

// Training schedule: the reference's HEAD schedule unless QB200_MODE selects an extension
// (1 = re-assign until convergence, 2 = that + empty-cell repair; include/qb200.h - parity unpinned).
int training_mode() {
  const char *e = std::getenv("QB200_MODE");
  const int m = e ? std::atoi(e) : QB200_MODE_PARITY;
  return (m == QB200_MODE_FULL || m == QB200_MODE_FULL_REPAIR) ? m : QB200_MODE_PARITY;
}

int checked_colorspace(ColorSpaces cs) {
  if (cs == ColorSpaces::NORMAL) return QB200_CS_NORMAL;
  if (cs == ColorSpaces::SCALED) return QB200_CS_SCALED;
  if (cs == ColorSpaces::CIE1931) return QB200_CS_CIE1931;  // FP64 vectors on the device (include/qb200.h)
  throw std::runtime_error("compress: unknown colour space");
}

}  // namespace

std::vector<Vector> getBlocksAsVectorsFromImage(const RGBImage &image, int w, int h, const ColorSpacePtr &cs) {
  const size_t xs = image.xSize, ys = image.ySize, npix = xs * ys, dim = (size_t)3 * w * h;
  std::vector<Vector> out(ceil_div(xs, w) * ceil_div(ys, h), Vector(dim, 0.0));
  for_each_block_pixel(xs, ys, w, h, [&](size_t vec, size_t elem, size_t pixel) {
    if (pixel >= npix) return;  // past the end of the buffer: stays 0.0
    const RGBDouble c = cs->RGBtoColorSpace(image.img[pixel]);
    for (int ch = 0; ch < 3; ch++) out[vec][elem + ch] = c[ch];
  });
  return out;
}

std::vector<CharVector> vectorsToCharVectorsColorSpaced(const std::vector<Vector> &vectors, const ColorSpacePtr &cs) {
  std::vector<CharVector> out;
  out.reserve(vectors.size());
  for (const Vector &v : vectors) {
    CharVector c(v.size());
    for (size_t p = 0; p + 2 < v.size(); p += 3) {
      const RGB rgb = cs->colorSpaceToRGB(RGBDouble{{v[p], v[p + 1], v[p + 2]}});
      c[p] = rgb[0];
      c[p + 1] = rgb[1];
      c[p + 2] = rgb[2];
    }
    out.push_back(c);
  }
  return out;
}

RGBImage getImageFromVectors(const std::vector<CharVector> &blocks, int xSize, int ySize, int w, int h) {
  RGBImage im;
  im.xSize = xSize;
  im.ySize = ySize;
  const size_t npix = (size_t)xSize * ySize;
  im.img.assign(npix, RGB{{0, 0, 0}});
  // same visiting order as the reference's loops, so that where wrapped blocks overlap the later write wins
  for_each_block_pixel(xSize, ySize, w, h, [&](size_t vec, size_t elem, size_t pixel) {
    if (pixel >= npix) return;
    const CharVector &b = blocks[vec];
    im.img[pixel] = RGB{{b[elem], b[elem + 1], b[elem + 2]}};
  });
  return im;
}

std::pair<CompressedImage, CompressionRaport> CompressedImage::compress(const RGBImage &image, Quantizers quantizer,
                                                                        ColorSpaces colorSpace, int blockWidth,
                                                                        int blockHeight, VectorType eps, int N) {
  const int cs = checked_colorspace(colorSpace);
  if (quantizer != Quantizers::LBG)
    throw std::runtime_error("compress: only Quantizers::LBG exists (the reference dereferences a null quantiser here)");
  if (image.img.empty()) throw std::out_of_range("compress: empty image");
  qb200_ctx *ctx = qbhost::context();
  const size_t dim = (size_t)3 * blockWidth * blockHeight, K = (size_t)1 << N;
  std::vector<double> cb(K * dim);
  double lbg_distortion = 0;
  CompressedImage res;

  // ---- the reference's timed region (src/Compressor.cpp:118-123): block extraction + quantize ----
  const auto t0 = std::chrono::system_clock::now();
  // the index vector (8 bytes per block) is allocated and first touched by a host thread while the image is uploaded
  // and the GPU trains
  const size_t n_blocks = (blockWidth > 0 && blockHeight > 0)   // (invalid shapes: qb200_set_image reports them below)
                              ? ceil_div((size_t)image.xSize, (size_t)blockWidth) * ceil_div((size_t)image.ySize, (size_t)blockHeight)
                              : 0;
  std::exception_ptr alloc_error;
  std::thread prefault([&] {
    try {
      res.assignedCodeVector.resize(n_blocks);
    } catch (...) {
      alloc_error = std::current_exception();
    }
  });
  int rc = qb200_set_image(ctx, reinterpret_cast<const uint8_t *>(image.img.data()), image.xSize, image.ySize, blockWidth,
                           blockHeight, cs, 1, 0);
  const char *what = "qb200_set_image";
  if (rc == QB200_OK) {
    rc = qb200_train(ctx, N, eps, training_mode(), 0, nullptr, nullptr, cb.data(), &lbg_distortion, nullptr);
    what = "qb200_train";
  }
  prefault.join();
  if (alloc_error) std::rethrow_exception(alloc_error);
  qbhost::check(rc, what);
  if (qb200_num_vectors(ctx) != n_blocks) res.assignedCodeVector.resize(qb200_num_vectors(ctx));
  qbhost::check(qb200_get_assign_u64(ctx, reinterpret_cast<uint64_t *>(res.assignedCodeVector.data())),
                "qb200_get_assign_u64");
  const auto t1 = std::chrono::system_clock::now();

  std::vector<uint8_t> cbb(K * dim);
  qbhost::check(qb200_codebook_to_bytes(cb.data(), K, (int)dim, cs, cbb.data()), "qb200_codebook_to_bytes");
  res.codeVectors.resize(K);
  for (size_t k = 0; k < K; k++) {
    CharVector c(dim);
    for (size_t d = 0; d < dim; d++) c[d] = (char)cbb[k * dim + d];
    res.codeVectors[k] = c;
  }
  res.xSize = image.xSize;
  res.ySize = image.ySize;
  res.blockWidth = blockWidth;
  res.blockHeight = blockHeight;
  res.colorSpace = colorSpace;
  res.quantizer = quantizer;

  // report distortion (src/Compressor.cpp:137-146): decode and compare pixels as signed chars - on the GPU
  double mse = 0;
  qbhost::check(qb200_decode(ctx, cbb.data(), (uint32_t)K, nullptr, &mse), "qb200_decode");
  CompressionRaport rap;
  rap.distortion = mse;
  rap.bitsPerPixel = (float)res.sizeInBits() / (float)((size_t)image.xSize * image.ySize);
  rap.uncompressedSize = image.sizeInBytes();
  rap.compressedSize = res.sizeInBits() / 8;
  rap.compressionTime = t1 - t0;
  return std::make_pair(std::move(res), rap);
}

RGBImage CompressedImage::decompress(const CompressedImage &c) {
  std::vector<CharVector> blocks;
  blocks.reserve(c.assignedCodeVector.size());
  for (size_t idx : c.assignedCodeVector) blocks.push_back(c.codeVectors.at(idx));
  return getImageFromVectors(blocks, (int)c.xSize, (int)c.ySize, (int)c.blockWidth, (int)c.blockHeight);
}

// Bit-packed size estimate (src/Compressor.cpp:174-182); the file itself stores whole bytes per index.
size_t CompressedImage::sizeInBits() {
  const size_t bits = index_bits(codeVectors.size()) * assignedCodeVector.size() +
                      blockWidth * blockHeight * codeVectors.size() * 8 * 3;
  return ceil_div(bits, 8) * 8;
}

// .quant container (src/Compressor.cpp:190-227):
//   "<bits> <colorSpace> <N> <xSize> <ySize> <blockW> <blockH>\n"  ASCII
//   K * dim codebook bytes
//   N indices, each the low ceil(bits/8) bytes of a little-endian size_t
void CompressedImage::saveToFile(const std::string &path) {
  std::ofstream out(path, std::ios::binary);
  if (!out) throw std::runtime_error("cannot write " + path);
  const size_t bits = index_bits(codeVectors.size());
  out << bits << " " << (int)colorSpace << " " << assignedCodeVector.size() << " " << xSize << " " << ySize << " "
      << blockWidth << " " << blockHeight << "\n";
  for (const CharVector &c : codeVectors) out.write(c.data(), (std::streamsize)c.size());
  const size_t bpi = ceil_div(bits, 8);
  std::vector<char> packed(assignedCodeVector.size() * bpi);
  for (size_t i = 0; i < assignedCodeVector.size(); i++)
    for (size_t b = 0; b < bpi; b++) packed[i * bpi + b] = (char)((assignedCodeVector[i] >> (8 * b)) & 0xff);
  out.write(packed.data(), (std::streamsize)packed.size());
}

// Packed container (extension): "QP1 <bits> <colorSpace> <N> <xSize> <ySize> <blockW> <blockH>\n", K * dim codebook
// bytes, then the N indices as one LSB-first bit stream of <bits> bits each, zero padded to a whole byte.
void CompressedImage::saveToFilePacked(const std::string &path) {
  std::ofstream out(path, std::ios::binary);
  if (!out) throw std::runtime_error("cannot write " + path);
  const size_t bits = index_bits(codeVectors.size()), n = assignedCodeVector.size();
  out << "QP1 " << bits << " " << (int)colorSpace << " " << n << " " << xSize << " " << ySize << " " << blockWidth << " "
      << blockHeight << "\n";
  for (const CharVector &c : codeVectors) out.write(c.data(), (std::streamsize)c.size());
  std::vector<unsigned char> stream(ceil_div(n * bits, 8), 0);
  for (size_t i = 0; i < n; i++)
    for (size_t b = 0; b < bits; b++)
      if ((assignedCodeVector[i] >> b) & 1) stream[(i * bits + b) >> 3] |= (unsigned char)(1u << ((i * bits + b) & 7));
  out.write(reinterpret_cast<const char *>(stream.data()), (std::streamsize)stream.size());
}

// Entropy-coded container (extension; `quant --entropy`):
//   "QH1 <bits> <colorSpace> <N> <xSize> <ySize> <blockW> <blockH> <stream bytes>\n", K * dim codebook bytes, K bytes
//   of code lengths (0: index never used), then the indices as canonical Huffman codes, MSB first, zero padded to a byte.
// Canonical: symbols ordered by (length, index); the first code of the shortest length is 0, each next code is the
// previous + 1, shifted left when the length grows - so the lengths alone define the code.
namespace {

constexpr int kMaxCodeLen = 32;

// Huffman code lengths for the given counts (0 for unused symbols), limited to kMaxCodeLen.
std::vector<unsigned char> huffman_lengths(const std::vector<unsigned long long> &count) {
  const size_t K = count.size();
  std::vector<unsigned char> len(K, 0);
  struct Node {
    unsigned long long w;
    int parent;
  };
  std::vector<Node> nodes;
  std::vector<int> leaf_of;  // node index -> symbol (leaves first)
  for (size_t k = 0; k < K; k++)
    if (count[k]) {
      nodes.push_back({count[k], -1});
      leaf_of.push_back((int)k);
    }
  const size_t n_leaves = nodes.size();
  if (n_leaves == 0) return len;
  if (n_leaves == 1) {
    len[(size_t)leaf_of[0]] = 1;
    return len;
  }
  // two-queue construction: leaves sorted by weight, internal nodes are created in non-decreasing weight order
  std::vector<int> order(n_leaves);
  for (size_t i = 0; i < n_leaves; i++) order[i] = (int)i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return nodes[(size_t)a].w < nodes[(size_t)b].w; });
  size_t qa = 0, qb = n_leaves;  // next unused leaf (in `order`), next unused internal node (index into nodes)
  auto take = [&]() {
    const bool leaf = qa < n_leaves && (qb >= nodes.size() || nodes[(size_t)order[qa]].w <= nodes[qb].w);
    return leaf ? order[qa++] : (int)qb++;
  };
  while ((n_leaves - qa) + (nodes.size() - qb) > 1) {
    const int a = take(), b = take();
    nodes.push_back({nodes[(size_t)a].w + nodes[(size_t)b].w, -1});
    nodes[(size_t)a].parent = nodes[(size_t)b].parent = (int)nodes.size() - 1;
  }
  std::vector<int> depth(nodes.size(), 0);
  for (size_t i = nodes.size() - 1; i-- > 0;) depth[i] = depth[(size_t)nodes[i].parent] + 1;  // parents have larger indices
  for (size_t i = 0; i < n_leaves; i++) len[(size_t)leaf_of[i]] = (unsigned char)std::min(depth[i], 255);
  // length limit: clamp, then lengthen the longest codes that are still short of the limit until Kraft's sum fits
  unsigned long long kraft = 0;  // in units of 2^-kMaxCodeLen
  for (size_t k = 0; k < K; k++)
    if (len[k]) {
      if (len[k] > kMaxCodeLen) len[k] = kMaxCodeLen;
      kraft += 1ull << (kMaxCodeLen - len[k]);
    }
  while (kraft > (1ull << kMaxCodeLen)) {
    size_t best = K;
    for (size_t k = 0; k < K; k++)
      if (len[k] && len[k] < kMaxCodeLen && (best == K || len[k] > len[best])) best = k;
    if (best == K) throw std::logic_error("huffman_lengths: cannot limit the code length");
    kraft -= 1ull << (kMaxCodeLen - len[best] - 1);
    len[best]++;
  }
  return len;
}

// canonical codes from lengths; first_code / first_pos / n_of per length serve the decoder
struct CanonicalCode {
  std::vector<unsigned int> code;     // per symbol
  std::vector<unsigned int> sorted;   // symbols by (length, index)
  unsigned long long first_code[kMaxCodeLen + 2];
  size_t first_pos[kMaxCodeLen + 2], n_of[kMaxCodeLen + 2];
};
CanonicalCode canonical_code(const std::vector<unsigned char> &len) {
  CanonicalCode c;
  c.code.assign(len.size(), 0);
  for (int l = 0; l <= kMaxCodeLen + 1; l++) c.first_code[l] = c.first_pos[l] = c.n_of[l] = 0;
  for (unsigned char l : len) {
    if (l > kMaxCodeLen) throw std::runtime_error("entropy container: code length out of range");
    c.n_of[l]++;
  }
  c.n_of[0] = 0;
  unsigned long long code = 0;
  size_t pos = 0;
  for (int l = 1; l <= kMaxCodeLen; l++) {
    code <<= 1;
    c.first_code[l] = code;
    c.first_pos[l] = pos;
    code += c.n_of[l];
    pos += c.n_of[l];
    if (code > (1ull << l)) throw std::runtime_error("entropy container: code lengths oversubscribed");
  }
  c.sorted.assign(pos, 0);
  std::vector<size_t> next(c.first_pos, c.first_pos + kMaxCodeLen + 2);
  for (size_t k = 0; k < len.size(); k++)
    if (len[k]) {
      c.code[k] = (unsigned int)(c.first_code[len[k]] + (next[len[k]] - c.first_pos[len[k]]));
      c.sorted[next[len[k]]++] = (unsigned int)k;
    }
  return c;
}

}  // namespace

void CompressedImage::saveToFileEntropy(const std::string &path) {
  std::ofstream out(path, std::ios::binary);
  if (!out) throw std::runtime_error("cannot write " + path);
  const size_t K = codeVectors.size(), bits = index_bits(K), n = assignedCodeVector.size();
  std::vector<unsigned long long> count(K, 0);
  for (size_t idx : assignedCodeVector) count.at(idx)++;
  const std::vector<unsigned char> len = huffman_lengths(count);
  const CanonicalCode cc = canonical_code(len);
  unsigned long long total_bits = 0;
  for (size_t k = 0; k < K; k++) total_bits += count[k] * len[k];
  std::vector<unsigned char> stream((size_t)((total_bits + 7) / 8), 0);
  unsigned long long at = 0;
  for (size_t idx : assignedCodeVector) {
    const unsigned int code = cc.code[idx];
    for (int b = len[idx] - 1; b >= 0; b--, at++)
      if ((code >> b) & 1u) stream[(size_t)(at >> 3)] |= (unsigned char)(0x80u >> (at & 7));
  }
  out << "QH1 " << bits << " " << (int)colorSpace << " " << n << " " << xSize << " " << ySize << " " << blockWidth << " "
      << blockHeight << " " << stream.size() << "\n";
  for (const CharVector &c : codeVectors) out.write(c.data(), (std::streamsize)c.size());
  out.write(reinterpret_cast<const char *>(len.data()), (std::streamsize)len.size());
  out.write(reinterpret_cast<const char *>(stream.data()), (std::streamsize)stream.size());
}

void CompressedImage::loadFromFile(const std::string &path) {
  std::ifstream in(path, std::ios::binary);
  if (!in) throw std::runtime_error("cannot open " + path);
  size_t bits = 0, n = 0;
  long long cs = 0;
  bool packed = false, entropy = false;
  size_t stream_bytes = 0;
  if (in.peek() == 'Q') {
    std::string magic;
    in >> magic;
    packed = magic == "QP1";
    entropy = magic == "QH1";
    if (!packed && !entropy) throw std::runtime_error(path + ": unknown container " + magic);
  }
  in >> bits >> cs >> n >> xSize >> ySize >> blockWidth >> blockHeight;
  if (entropy) in >> stream_bytes;
  if (!in || bits > 24 || blockWidth == 0 || blockHeight == 0) throw std::runtime_error(path + ": bad .quant header");
  in.get();  // '\n'
  // files written by the reference carry an uninitialised value in this field; decoding never uses it
  colorSpace = (cs >= 0 && cs <= 2) ? (ColorSpaces)cs : ColorSpaces::SCALED;
  const size_t K = (size_t)1 << bits, dim = blockWidth * blockHeight * 3, bpi = ceil_div(bits, 8);
  codeVectors.assign(K, CharVector(dim));
  for (CharVector &c : codeVectors) in.read(c.data(), (std::streamsize)dim);
  if (entropy) {
    std::vector<unsigned char> len(K), stream(stream_bytes);
    in.read(reinterpret_cast<char *>(len.data()), (std::streamsize)K);
    in.read(reinterpret_cast<char *>(stream.data()), (std::streamsize)stream_bytes);
    if (!in) throw std::runtime_error(path + ": truncated .quant file");
    const CanonicalCode cc = canonical_code(len);
    assignedCodeVector.assign(n, 0);
    const unsigned long long avail = (unsigned long long)stream_bytes * 8;
    unsigned long long at = 0;
    for (size_t i = 0; i < n; i++) {
      unsigned long long code = 0;
      int l = 0;
      for (;;) {
        if (at >= avail || l >= kMaxCodeLen) throw std::runtime_error(path + ": corrupt entropy-coded index stream");
        code = (code << 1) | ((stream[(size_t)(at >> 3)] >> (7 - (at & 7))) & 1u);
        at++;
        l++;
        if (cc.n_of[l] && code >= cc.first_code[l] && code - cc.first_code[l] < cc.n_of[l]) break;
      }
      assignedCodeVector[i] = cc.sorted[cc.first_pos[l] + (size_t)(code - cc.first_code[l])];
    }
    return;
  }
  std::vector<unsigned char> raw(packed ? ceil_div(n * bits, 8) : n * bpi);
  in.read(reinterpret_cast<char *>(raw.data()), (std::streamsize)raw.size());
  if (!in) throw std::runtime_error(path + ": truncated .quant file");
  assignedCodeVector.assign(n, 0);
  if (packed) {
    for (size_t i = 0; i < n; i++)
      for (size_t b = 0; b < bits; b++)
        assignedCodeVector[i] |= (size_t)((raw[(i * bits + b) >> 3] >> ((i * bits + b) & 7)) & 1) << b;
    return;
  }
  for (size_t i = 0; i < n; i++)
    for (size_t b = 0; b < bpi; b++) assignedCodeVector[i] |= (size_t)raw[i * bpi + b] << (8 * b);
}

std::ostream &operator<<(std::ostream &s, const CompressionRaport &r) {
  s << "Compression raport: " << std::endl;
  s << "Distortion        = " << std::fixed << std::setprecision(10) << r.distortion << std::endl;
  s << "Bits per pixel    = " << r.bitsPerPixel << std::endl;
  s << "Uncompressed size = " << pretty_bytes(r.uncompressedSize) << std::endl;
  s << "Compressed size   = " << pretty_bytes(r.compressedSize) << std::endl;
  s << "Compression ratio = " << std::fixed << std::setprecision(3) << (double)r.compressedSize / r.uncompressedSize
    << std::endl;
  s << "Compression time  = " << r.compressionTime.count() << "s" << std::endl;
  return s;
}
