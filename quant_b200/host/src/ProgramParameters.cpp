#include "ProgramParameters.hpp"

#include "ColorSpace.hpp"
#include "Quantizer.hpp"

static ProgramParameters g_params;

ProgramParameters *getParams() { return &g_params; }

// Defaults of the reference's CLI (src/main.cpp:49-57).
void paramsInitialize() {
  g_params.n = 8;
  g_params.width = 2;
  g_params.height = 2;
  g_params.eps = 0.000001f;
  g_params.raport = false;
  g_params.show = false;
  g_params.quantizer = (int)Quantizers::LBG;
  g_params.colorspace = (int)ColorSpaces::SCALED;
  g_params.pack = false;
  g_params.entropy = false;
  g_params.file.clear();
  g_params.saveto.clear();
}
