// oracle/ref_nodump.cpp — LD_PRELOAD no-op for the reference's dump_mesh() (source/kernel.cpp:158-186),
// which is called unconditionally (preproc.cpp:2371 hard-codes verbose=true) and writes ~17 .off files
// into the CWD per mcDispatch.  TEST/BENCH INFRASTRUCTURE ONLY; changes no result.
class hmesh_t;
void dump_mesh(const hmesh_t&, const char*, const double) { }
