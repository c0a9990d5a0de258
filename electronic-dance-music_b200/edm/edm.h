// EDM::edm_error — the reference's fatal-error convention (lib/edm.h:1-7, lib/edm.cpp:4-7):
// message + location on stderr, then abort().  The C ABI underneath returns status codes; the C++
// mirror turns any failure into this call so callers see the reference's behaviour.
#ifndef EDM_B200_EDM_H_
#define EDM_B200_EDM_H_

#include <iostream>

namespace EDM {

void edm_error(const char* error, const char* location);

// aborts through edm_error when a C-ABI call failed (status != 0)
void edm_check(int status, const char* location);

}  // namespace EDM

#endif
