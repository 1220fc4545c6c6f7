// TEST INFRASTRUCTURE ONLY: exposes the reference's own hash (dysgu/include/xxhash64.h, bound by map_set_utils.pxd:28-29) with C
// linkage.  Compiled by oracle/build_ref.sh with -I<reference>/dysgu; the header itself is not copied.
#include <stdint.h>
#include "include/xxhash64.h"
extern "C" uint64_t ref_xxh64(const void *input, uint64_t length, uint64_t seed) { return XXHash64::hash(input, length, seed); }
