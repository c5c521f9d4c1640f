/* Drop-in include path: code written against the reference (`#include "omega/list_matcher.h"`,
 * e.g. omega_match/main.c:30) compiles against the B200 library through this shim. */
#ifndef OMEGA_LIST_MATCHER_H
#define OMEGA_LIST_MATCHER_H
#include "../olm_b200.h"
#endif
