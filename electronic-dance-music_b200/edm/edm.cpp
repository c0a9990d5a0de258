#include "edm.h"

#include <cstdlib>

#include "../../include/edm_b200.h"

namespace EDM {

void edm_error(const char* error, const char* location) {
  std::cerr << "[EDM:" << location << "] " << error << std::endl;
  abort();
}

void edm_check(int status, const char* location) {
  if (status != EDM_OK) edm_error(edm_last_error(), location);
}

}  // namespace EDM
