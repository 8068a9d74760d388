/* <huffman/sys.h> — include-path and source compatibility with the reference's header of this
 * name [ref: include/huffman/sys.h:9-74].  All declarations live in <huffman.h>; this header
 * adds the reference's public "routine" macro vocabulary (a function keeps its status in a
 * local `__error` and funnels every exit through the label `ensure:`), so that caller code
 * written against those macros still compiles.  The library itself does not use them
 * (csrc/host/internal.h has its own early-return helpers).
 *
 *   huf_error_t f(T *p) {
 *       routine_m();                  // declares __error = HUF_ERROR_SUCCESS
 *       routine_param_m(p);           // p == 0 -> INVALID_ARGUMENT, jumps to the label
 *       ...  routine_error_m(e);      // leave with e       routine_success_m();  // leave with OK
 *       routine_ensure_m();           // the label; cleanup that must always run follows it
 *       ...  if (routine_violation_m()) { undo }
 *       routine_defer_m();            // return __error
 *   }
 */
#ifndef HUFFMAN_B200_SYS_H
#define HUFFMAN_B200_SYS_H

#include <stdio.h>

#include "../huffman.h"

/* cast for the void** out-parameters of huf_malloc [ref: sys.h:9] */
#define void_pptr_m(pointer) ((void **)(pointer))

/* status variable of the routine [ref: sys.h:13-14] */
#define routine_m() huf_error_t __error = HUF_ERROR_SUCCESS

/* the finalisation label [ref: sys.h:18-19] */
#define routine_ensure_m() ensure:

/* return the status; placed behind the finalisation code [ref: sys.h:23-26] */
#define routine_defer_m() \
    do {                  \
        return __error;   \
    } while (0)

/* label and return in one, for routines without finalisation code [ref: sys.h:30-34] */
#define routine_yield_m()   \
    do {                    \
        routine_ensure_m(); \
        return __error;     \
    } while (0)

/* a nil parameter ends the routine with HUF_ERROR_INVALID_ARGUMENT [ref: sys.h:38-44] */
#define routine_param_m(param)                    \
    do {                                          \
        if ((param) == 0) {                       \
            __error = HUF_ERROR_INVALID_ARGUMENT; \
            goto ensure;                          \
        }                                         \
    } while (0)

/* so does a value outside [low, high] [ref: sys.h:49-55] */
#define routine_inrange_m(value, low, high)           \
    do {                                              \
        if ((value) < (low) || (value) > (high)) {    \
            __error = HUF_ERROR_INVALID_ARGUMENT;     \
            goto ensure;                              \
        }                                             \
    } while (0)

/* leave with the given status [ref: sys.h:59-63] */
#define routine_error_m(error) \
    do {                       \
        __error = (error);     \
        goto ensure;           \
    } while (0)

/* leave with HUF_ERROR_SUCCESS [ref: sys.h:67-68] */
#define routine_success_m() routine_error_m(HUF_ERROR_SUCCESS)

/* true once the routine was left with an error [ref: sys.h:73-74] */
#define routine_violation_m() (__error != HUF_ERROR_SUCCESS)

#endif /* HUFFMAN_B200_SYS_H */
