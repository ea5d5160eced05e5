# cython: language_level=3
# Drop-in replacement of the reference's Cython module `nem` (ppanggolin/NEM/nem.pyx:1-14,
# built by the reference's setup.py:51-59; imported with `from nem import *` at
# ppanggolin/ppanggolin.py:20 and called with keywords at ppanggolin.py:1814-1826).
# Same module name, same function name, same 13 keyword arguments (bytes for the strings), same
# int return value (ExitET, NEM/lib_io.h:22-34) -- but the symbol it binds is the B200 engine's
# nem() in libnem_b200.so (include/nem_b200.h) instead of NEM/nem_exe.c.
cdef extern from "nem_b200.h":
    int c_nem "nem"(const char *Fname, const int nk, const char *algo, const float beta,
                    const char *convergence, const float convergence_th, const char *format,
                    const int it_max, const int dolog, const char *model_family,
                    const char *proportion, const char *dispersion, const int init_mode) nogil


def nem(bytes Fname, int nk, bytes algo, float beta, bytes convergence, float convergence_th,
        bytes format, int it_max, bint dolog, bytes model_family, bytes proportion,
        bytes dispersion, int init_mode):
    cdef const char *c_fname = Fname
    cdef const char *c_algo = algo
    cdef const char *c_conv = convergence
    cdef const char *c_fmt = format
    cdef const char *c_fam = model_family
    cdef const char *c_prop = proportion
    cdef const char *c_disp = dispersion
    cdef int rc
    with nogil:     # the fit blocks on the GPU; let other Python threads run
        rc = c_nem(c_fname, nk, c_algo, beta, c_conv, convergence_th, c_fmt, it_max, dolog,
                   c_fam, c_prop, c_disp, init_mode)
    return rc
