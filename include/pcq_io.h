/*
 * pcq_io.h — C-ABI of libpcq_io: the on-disk sample format either side of the PC hot path
 * (second half).
 *
 * PyTUQ's applications exchange germ samples, parameter samples and model outputs as
 * whitespace-separated text written by numpy.savetxt with its defaults (fmt '%.18e',
 * delimiter ' ', newline '\n') and read back with numpy.loadtxt
 * (apps/uqpc/uq_pc.py:202-271, the other apps and examples).  At 1e7 rows this
 * text I/O, not the fit, dominates wall time.  This library is the same format, byte for
 * byte on write and bit for bit on read, done by all host cores:
 *
 *   - writing formats blocks of rows in parallel (each value with the C locale's
 *     "%.18e", which is the correctly rounded 19-digit string Python's '%.18e' % v
 *     produces) and appends the blocks to the file in order;
 *   - reading maps the file, cuts it at line starts into segments, counts the data rows of
 *     every segment, then parses the segments in parallel straight into the caller's
 *     (rows x cols) float64 buffer with a correctly rounded decimal -> binary64 conversion
 *     (strtod in the C locale), i.e. the value numpy.loadtxt returns for the same token.
 *
 * Host-only (no CUDA): it lives in its own shared library so that it can be used on boxes
 * without a GPU (the black-box model step of uq_pc.py) and cannot disturb libpcq.
 *
 * Conventions as in pcq.h: plain pointers and sizes, 0 on success, negative status
 * otherwise, never throws, pcqio_last_error() gives a thread-local description.
 */
#ifndef PCQ_IO_H_
#define PCQ_IO_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCQIO_VERSION 100

typedef enum pcqio_status {
    PCQIO_OK = 0,
    PCQIO_ERR_ARG = -1,     /* null pointer, negative size, ld < cols */
    PCQIO_ERR_IO = -2,      /* open / map / write failed (errno text in pcqio_last_error) */
    PCQIO_ERR_PARSE = -3,   /* a token is not a number numpy.loadtxt would accept */
    PCQIO_ERR_RAGGED = -4,  /* rows with different numbers of columns */
    PCQIO_ERR_ALLOC = -5
} pcqio_status;

typedef struct pcqio_reader pcqio_reader;

int pcqio_version(void);
const char* pcqio_last_error(void);

/* numpy.savetxt(path, a) with numpy's defaults, for a rows x cols row-major float64 matrix with
 * leading dimension ld >= cols (a 1-D array is written as rows x 1, as numpy does:
 * lib/_npyio_impl.py, savetxt).  threads <= 0: all online cores.
 * Replaces np.savetxt at apps/uqpc/uq_pc.py:202-228,252-258. */
int pcqio_savetxt(const char* path, const double* a, int64_t rows, int64_t cols, int64_t ld,
                  int threads);

/* numpy.loadtxt(path) with numpy's defaults (whitespace-separated, '#' starts a comment, blank
 * lines skipped), in two steps so that the caller owns the result buffer:
 *   pcqio_open   maps the file and counts: *rows data rows of *cols values each
 *                (0 x 0 for a file without data rows; PCQIO_ERR_RAGGED if the first rows of two
 *                segments disagree — every row is checked again while parsing);
 *   pcqio_read   parses into out (rows x cols, leading dimension ld >= cols);
 *   pcqio_close  unmaps.
 * Replaces np.loadtxt at apps/uqpc/uq_pc.py:232-271. */
int pcqio_open(const char* path, pcqio_reader** reader, int64_t* rows, int64_t* cols, int threads);
int pcqio_read(pcqio_reader* reader, double* out, int64_t ld);
int pcqio_close(pcqio_reader* reader);

/* Conversions.  Both directions take one 64 x 128-bit product with a truncated power of ten
 * (tools/gen_pow10.py) and hand the rare value whose rounding that product cannot decide (and subnormal
 * results, inputs of more than 19 digits, inf / nan) to libc's correctly rounded snprintf / strtod in the C
 * locale.  pcqio_selftest converts n pseudo-random values (all bit patterns, short dyadic and decimal
 * fractions, powers of ten, subnormals, range edges) both ways and returns how many results differ from
 * libc's (0 expected); *slow_formats / *slow_parses receive how many conversions took the libc route. */
int64_t pcqio_selftest(uint64_t seed, int64_t n, int64_t* slow_formats, int64_t* slow_parses);

#ifdef __cplusplus
}
#endif
#endif /* PCQ_IO_H_ */
