#include "B200Context.hpp"

#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

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
    int rc;
    // QB200_DEVICES=all | 0,1,2,3 : shard every image over these GPUs (qb200_create_multi; the devices reduce the
    // per-level statistics among themselves over NVLink).  Otherwise one GPU: QB200_DEVICE (default 0).
    if (const char *list = std::getenv("QB200_DEVICES")) {
      std::vector<int> ids;
      const std::string text(list);
      if (text != "all") {
        size_t pos = 0;
        while (pos < text.size()) {
          size_t next = text.find(',', pos);
          if (next == std::string::npos) next = text.size();
          if (next > pos) ids.push_back(std::atoi(text.substr(pos, next - pos).c_str()));
          pos = next + 1;
        }
      }
      rc = qb200_create_multi((int)ids.size(), ids.empty() ? nullptr : ids.data(), &holder.ctx);
    } else {
      const char *dev = std::getenv("QB200_DEVICE");
      rc = qb200_create(dev ? std::atoi(dev) : 0, &holder.ctx);
    }
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
