#include "B200Context.hpp"

#include <cstdlib>
#include <stdexcept>
#include <string>

#include "../../../include/qb200.h"

namespace qbhost {

namespace {
struct Holder {
  qb200_ctx *ctx = nullptr;
  ~Holder() {
    if (ctx) qb200_destroy(ctx);
  }
};
thread_local Holder holder;
}  // namespace

qb200_ctx *context() {
  if (!holder.ctx) {
    const char *dev = std::getenv("QB200_DEVICE");
    const int rc = qb200_create(dev ? std::atoi(dev) : 0, &holder.ctx);
    if (rc != QB200_OK)
      throw std::runtime_error(std::string("quant (B200): ") + qb200_last_error(nullptr) +
                               " - this build has no CPU path");
  }
  return holder.ctx;
}

void check(int status, const char *what) {
  if (status == QB200_OK) return;
  const char *msg = holder.ctx ? qb200_last_error(holder.ctx) : "";
  throw std::runtime_error(std::string(what) + ": libqb200 error " + std::to_string(status) + ": " + msg);
}

}  // namespace qbhost
