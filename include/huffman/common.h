/* <huffman/common.h> — include-path compatibility with the reference's header of this name
 * [ref: include/huffman/common.h].  All declarations live in <huffman.h>. */
#include "../huffman.h"
