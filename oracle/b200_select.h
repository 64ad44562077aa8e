/*
 * oracle/b200_select.h -- TEST INFRASTRUCTURE.  Force-included (-include) when the reference's OWN,
 * unmodified stage wrappers are compiled against the B200 backend (Makefile target ref_b200):
 * it pre-defines the include guard of /root/reference/src/backend.h, so that header's compile-time
 * dispatch (src/backend.h:26-46) is skipped, and aliases paris::backend to paris::b200 instead --
 * exactly what the three-line patch in INTEGRATION.md does inside the reference tree.
 */
#ifndef PARIS_BACKEND_H_
#define PARIS_BACKEND_H_

#include "../paris_b200/cpp/b200/backend.h"

namespace paris
{
    namespace backend = b200;
}

#endif
