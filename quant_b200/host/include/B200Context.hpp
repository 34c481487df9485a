// Process-wide libqb200 context shared by Quantizer.cpp and Compressor.cpp (one per host thread).
#pragma once
struct qb200_ctx;
namespace qbhost {
// Creates the context on first use (device = $QB200_DEVICE or 0); throws std::runtime_error when no
// B200 is usable.  check() turns a libqb200 status into std::runtime_error with the library's text.
qb200_ctx *context();
void check(int status, const char *what);
}  // namespace qbhost
