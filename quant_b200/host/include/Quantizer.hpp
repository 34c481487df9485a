// The quantiser plugin interface of the reference (/root/reference/include/Quantizer.hpp:8-20),
// unchanged: same enum, same abstract class, same factory.  getQuantizer(Quantizers::LBG) returns
// a quantiser whose quantize() runs on a B200 through libqb200 (include/qb200.h); the other enum
// values return nullptr exactly as the reference does (src/Quantizer.cpp:146-155).
//
//   codebook, assignment, distortion = getQuantizer(Quantizers::LBG)->quantize(trainingSet, n, eps)
//
// trainingSet: N vectors of equal dimension; n: bits, K = 2^n codevectors; eps: the reference's
// convergence threshold (it has no effect on the HEAD schedule).
// Errors: std::out_of_range for an empty training set (the reference's trainingSet.at(0));
// std::runtime_error if no B200 is usable - there is no CPU fallback.  Vectors on the NORMAL / SCALED
// byte lattice (what getBlocksAsVectorsFromImage produces) train on the fast integer path, any other
// doubles (CIE1931, arbitrary data; dimension <= 192) as FP64 vectors - same results, slower.
#pragma once
#include <memory>
#include <tuple>
#include <vector>

#include "VectorOperations.hpp"

enum class Quantizers { LBG, MEDIAN_CUT, LBG_MEDIAN_CUT, ABC };

class AbstractQuantizer {
 public:
  virtual std::tuple<std::vector<Vector>, std::vector<size_t>, VectorType> quantize(
      const std::vector<Vector> &trainingSet, size_t n, VectorType eps) = 0;
  virtual ~AbstractQuantizer() = default;
};

typedef std::unique_ptr<AbstractQuantizer> QuantizerPtr;

QuantizerPtr getQuantizer(Quantizers);
