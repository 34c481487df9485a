// quant - command-line front end with the reference's flags (/root/reference/src/main.cpp:45-112),
// parsed without boost::program_options:
//   quant <file> -o <saveto> [-n bits] [-e eps] [-w W] [-h H] [-r 1] [-q quantizer] [--c colorspace] [--help]
//   plus one extension: --pack writes the bit-packed .quant container (reading detects either container).
// Mode by file extension: .ppm -> .quant compresses, .quant -> .ppm decompresses, .ppm -> .ppm does both.
#include <cstdlib>
#include <iostream>
#include <stdexcept>
#include <string>

#include "Compressor.hpp"
#include "ProgramParameters.hpp"

namespace {

enum class FileType { PPM, QUANT, OTHER };

FileType file_type(const std::string &p) {
  auto ends = [&](const char *suf) {
    const std::string s(suf);
    return p.size() >= s.size() && p.compare(p.size() - s.size(), s.size(), s) == 0;
  };
  if (ends(".ppm")) return FileType::PPM;
  if (ends(".quant")) return FileType::QUANT;
  return FileType::OTHER;
}

void usage() {
  std::cout << "Options:\n"
               "  --help                 Print help messages\n"
               "  -n arg (=8)            bits per codevector\n"
               "  -e arg (=1e-06)        eps parameter for quantization algorithm\n"
               "  -w arg (=2)            Width of block\n"
               "  -h arg (=2)            Height of block\n"
               "  --file arg             File to compress/decompress\n"
               "  -o [ --saveto ] arg    Save to\n"
               "  -r arg (=0)            Print raport to std::out\n"
               "  -q [ --quantizer ] arg (=0)  Pick quantizer\n"
               "  --c [ --colorspace ] arg (=1) Pick ColorSpace\n"
               "  --pack                 (extension) bit-packed indices in the .quant file\n"
               "  --entropy              (extension) Huffman-coded indices in the .quant file\n"
               "                         (.quant -> .quant re-writes a file in the chosen container)\n";
}

bool parse_bool(const std::string &v) { return v == "1" || v == "true" || v == "yes" || v == "on"; }

}  // namespace

int main(int argc, char **argv) {
  paramsInitialize();
  ProgramParameters *par = getParams();
  try {
    for (int i = 1; i < argc; i++) {
      const std::string a = argv[i];
      auto value = [&]() -> std::string {
        if (i + 1 >= argc) throw std::runtime_error("the required argument for option '" + a + "' is missing");
        return argv[++i];
      };
      if (a == "--help") { usage(); return 0; }
      else if (a == "-n") par->n = std::stoi(value());
      else if (a == "-e") par->eps = std::stof(value());
      else if (a == "-w") par->width = std::stoi(value());
      else if (a == "-h") par->height = std::stoi(value());
      else if (a == "-o" || a == "--saveto") par->saveto = value();
      else if (a == "--file") par->file = value();
      else if (a == "-r") par->raport = (i + 1 < argc && argv[i + 1][0] != '-') ? parse_bool(value()) : true;
      else if (a == "-q" || a == "--quantizer") par->quantizer = std::stoi(value());
      else if (a == "--c" || a == "--colorspace") par->colorspace = std::stoi(value());
      else if (a == "--pack") par->pack = true;
      else if (a == "--entropy") par->entropy = true;
      else if (!a.empty() && a[0] == '-') throw std::runtime_error("unrecognised option '" + a + "'");
      else par->file = a;  // positional: the input file
    }
    if (par->file.empty()) throw std::runtime_error("the option '--file' is required but missing");
    if (par->saveto.empty()) throw std::runtime_error("the option '--saveto' is required but missing");
    if (par->n < 0 || par->n > 16) throw std::runtime_error("-n must be in [0, 16]");

    const FileType from = file_type(par->file), to = file_type(par->saveto);
    auto run_compression = [&]() {
      RGBImage img(par->file);
      auto result = CompressedImage::compress(img, (Quantizers)par->quantizer, (ColorSpaces)par->colorspace, par->width,
                                              par->height, par->eps, par->n);
      if (par->raport) std::cout << result.second;
      return result.first;
    };
    if (from == FileType::PPM && to == FileType::PPM) {
      CompressedImage c = run_compression();
      CompressedImage::decompress(c).saveToFile(par->saveto);
    } else if (from == FileType::QUANT && to == FileType::PPM) {
      CompressedImage c;
      c.loadFromFile(par->file);
      CompressedImage::decompress(c).saveToFile(par->saveto);
    } else if (from == FileType::PPM && to == FileType::QUANT) {
      CompressedImage c = run_compression();
      if (par->entropy) c.saveToFileEntropy(par->saveto);
      else if (par->pack) c.saveToFilePacked(par->saveto);
      else c.saveToFile(par->saveto);
    } else if (from == FileType::QUANT && to == FileType::QUANT) {  // (extension) re-write in another container: host only
      CompressedImage c;
      c.loadFromFile(par->file);
      if (par->entropy) c.saveToFileEntropy(par->saveto);
      else if (par->pack) c.saveToFilePacked(par->saveto);
      else c.saveToFile(par->saveto);
    } else {
      std::cerr << "File type not supported" << std::endl;
      return 1;
    }
  } catch (const std::exception &e) {
    std::cerr << "quant: " << e.what() << std::endl;
    return 2;
  }
  return 0;
}
