// Process-wide state of the C++ drop-in wrappers (myLSD_b200.cpp, myFA_b200.cpp).
//
// The reference's entry points (LSD/myLSD.h:131-132, LSD/myFA.h:83) carry no context argument, so the
// wrappers share one lazily created lsdb_ctx on the device named by $LSDB_DEVICE (default 0).
// Failures cannot be reported through the reference's signatures ("never fails" semantics, SURVEY.md
// §8b): the wrappers print the library's error text and abort — there is no CPU fallback.
#pragma once
#include "../../include/lsdb200.h"

namespace lsdb_host {
lsdb_ctx* context();                                 // creates the context on first use; aborts on failure
[[noreturn]] void die(const char* what, int rc);     // prints lsdb_last_error and aborts
void shutdown();                                     // optional: releases the context and cached device buffers
}  // namespace lsdb_host
