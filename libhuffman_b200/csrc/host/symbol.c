/*
 * symbol.c — byte -> code-string table objects [ref: src/symbol.c:10-222].
 * Host-side objects for API/link compatibility (test/symbol_test.c); on the GPU the table is
 * a 256-entry {code,len} array built by kernel K2.
 */
#include <string.h>

#include "internal.h"

huf_error_t
huf_symbol_mapping_element_init(huf_symbol_mapping_element_t **self, const uint8_t *coding,
                                size_t length)
{
    HUF_REQUIRE(self);
    HUF_REQUIRE(coding);
    HUF_TRY(huf_malloc((void **)self, sizeof(**self), 1));
    /* one spare byte keeps the copy NUL terminated */
    HUF_TRY(huf_malloc((void **)&(*self)->coding, 1, length + 1));
    memcpy((*self)->coding, coding, length);
    (*self)->length = length;
    return HUF_ERROR_SUCCESS;
}

huf_error_t
huf_symbol_mapping_element_free(huf_symbol_mapping_element_t **self)
{
    HUF_REQUIRE(self);
    if (*self) {
        free((*self)->coding);
        free(*self);
    }
    *self = NULL;
    return HUF_ERROR_SUCCESS;
}

huf_error_t
huf_symbol_mapping_init(huf_symbol_mapping_t **self, size_t length)
{
    HUF_REQUIRE(self);
    HUF_TRY(huf_malloc((void **)self, sizeof(**self), 1));
    HUF_TRY(huf_malloc((void **)&(*self)->symbols, sizeof(void *), length));
    (*self)->length = length;
    return HUF_ERROR_SUCCESS;
}

static void
drop_all(huf_symbol_mapping_t *m)
{
    for (size_t i = 0; i < m->length; i++) {
        if (m->symbols[i]) {
            huf_symbol_mapping_element_free(&m->symbols[i]);
        }
    }
}

huf_error_t
huf_symbol_mapping_free(huf_symbol_mapping_t **self)
{
    HUF_REQUIRE(self);
    if (*self) {
        drop_all(*self);
        free((*self)->symbols);
        free(*self);
    }
    *self = NULL;
    return HUF_ERROR_SUCCESS;
}

huf_error_t
huf_symbol_mapping_reset(huf_symbol_mapping_t *self)
{
    HUF_REQUIRE(self);
    drop_all(self);
    return HUF_ERROR_SUCCESS;
}

/* Takes ownership of `element`; a previous occupant of the slot is released. */
huf_error_t
huf_symbol_mapping_insert(huf_symbol_mapping_t *self, size_t position,
                          huf_symbol_mapping_element_t *element)
{
    HUF_REQUIRE(self);
    HUF_REQUIRE(element);
    if (position >= self->length) {
        return HUF_ERROR_INVALID_ARGUMENT;
    }
    if (self->symbols[position]) {
        huf_symbol_mapping_element_free(&self->symbols[position]);
    }
    self->symbols[position] = element;
    return HUF_ERROR_SUCCESS;
}

huf_error_t
huf_symbol_mapping_get(huf_symbol_mapping_t *self, size_t position,
                       huf_symbol_mapping_element_t **element)
{
    HUF_REQUIRE(self);
    HUF_REQUIRE(element);
    if (position >= self->length) {
        return HUF_ERROR_INVALID_ARGUMENT;
    }
    *element = self->symbols[position];
    return HUF_ERROR_SUCCESS;
}
