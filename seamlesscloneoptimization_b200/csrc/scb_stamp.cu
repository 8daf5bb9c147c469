// scb_stamp.cu -- the build stamp: a hash of every source file of the library (__graft_entry__.source_hash), passed by the
// build as -DSCB_SOURCE_HASH_STR.  tests/conftest.py compares it with the tree and rebuilds a stale library, so a parity run
// can never use a libscb.so that does not match the reviewed sources.  Kept in its own translation unit so that editing one
// kernel file recompiles that file and this one only.
#ifndef SCB_SOURCE_HASH_STR
#define SCB_SOURCE_HASH_STR "unstamped-build!"
#endif
extern "C" const char* scb_source_hash(void) {
    static const char stamp[] = "SCB_SOURCE_HASH=" SCB_SOURCE_HASH_STR;
    return stamp + 16;
}
