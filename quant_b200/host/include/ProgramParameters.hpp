// Command-line parameters of the `quant` tool, shared through getParams() exactly like the reference's global
// (/root/reference/include/ProgramParameters.hpp:5-19, filled by src/main.cpp:45-57).  Field names are the
// reference's - client code reads them by name - with the CLI flag and default each one carries.
#pragma once
#include <string>

struct ProgramParameters {
  // --- what to do -------------------------------------------------------------------------------------------
  std::string file;    // positional / --file : input, .ppm (compress) or .quant (decompress)
  std::string saveto;  // -o / --saveto       : output, .quant or .ppm
  bool raport;         // -r                  : print the compression report to stdout       (default false)
  bool show;           // unused by the reference's main(); kept for layout compatibility of client code

  // --- how to quantise --------------------------------------------------------------------------------------
  int n;               // -n : bits per index, K = 2^n codevectors                             (default 8)
  float eps;           // -e : LBG convergence threshold (float, as in the reference)         (default 1e-6)
  int width;           // -w : block width  in pixels                                         (default 2)
  int height;          // -h : block height in pixels                                         (default 2)
  int quantizer;       // -q / --quantizer  : enum Quantizers as int                           (default LBG = 0)
  int colorspace;      // --c / --colorspace : enum ColorSpaces as int                         (default SCALED = 1)

  // --- extension, not in the reference ---------------------------------------------------------------------
  bool pack;           // --pack : write the bit-packed .quant container (Compressor.hpp)      (default false)
  bool entropy;        // --entropy : write the Huffman-coded .quant container                 (default false)
};

// The process-wide instance, and a reset to the defaults listed above.
ProgramParameters *getParams();
void paramsInitialize();
