/* errors.c — huf_error_string [ref: src/errors.c:5-33]. */
#include "internal.h"

const char *
huf_error_string(huf_error_t error)
{
    switch ((int)error) {
    case HUF_ERROR_SUCCESS:
        return "Success";
    case HUF_ERROR_MEMORY_ALLOCATION:
        return "Failed to allocate the requested memory block";
    case HUF_ERROR_INVALID_ARGUMENT:
        return "An invalid argument was specified to the function";
    case HUF_ERROR_READ_WRITE:
        return "Failed on read/write operation";
    case HUF_ERROR_FATAL:
        return "Fatal error";
    case HUF_ERROR_BTREE_OVERFLOW:
        return "Block is corrupted, Huffman tree has impossible size";
    case HUF_ERROR_BTREE_CORRUPTED:
        return "Huffman tree is corrupted and cannot be used to decode the block";
    default:
        return "Unknown error";
    }
}
