/* Minimal stand-in for libogg's <ogg/ogg.h>, written for this repo.
 *
 * TEST INFRASTRUCTURE ONLY.  libogg (xiph/ogg, "ogg >= 1.0" per the reference's
 * configure.ac:247) is an external dependency of libvorbis that is not
 * installed in this image.  libvorbis needs only integer typedefs, the
 * allocator macros, two plain structs and eleven bit-packer functions from
 * it.  This header declares exactly those so the
 * reference C sources compile unmodified into oracle/_ref/.
 * The bit-packer itself (LSb-first, doc/02-bitpacking.tex) is bitpack.c.
 */
#ifndef VB200_OGG_SHIM_H
#define VB200_OGG_SHIM_H
#include <stdlib.h>
#include <stdint.h>

typedef int16_t  ogg_int16_t;
typedef uint16_t ogg_uint16_t;
typedef int32_t  ogg_int32_t;
typedef uint32_t ogg_uint32_t;
typedef int64_t  ogg_int64_t;
typedef uint64_t ogg_uint64_t;

#define _ogg_malloc  malloc
#define _ogg_calloc  calloc
#define _ogg_realloc realloc
#define _ogg_free    free

typedef struct {
  long endbyte;
  int  endbit;
  unsigned char *buffer;
  unsigned char *ptr;
  long storage;
} oggpack_buffer;

typedef struct {
  unsigned char *packet;
  long  bytes;
  long  b_o_s;
  long  e_o_s;
  ogg_int64_t granulepos;
  ogg_int64_t packetno;
} ogg_packet;

void  oggpack_writeinit(oggpack_buffer *b);
void  oggpack_reset(oggpack_buffer *b);
void  oggpack_writeclear(oggpack_buffer *b);
void  oggpack_writetrunc(oggpack_buffer *b, long bits);
void  oggpack_write(oggpack_buffer *b, unsigned long value, int bits);
void  oggpack_readinit(oggpack_buffer *b, unsigned char *buf, int bytes);
long  oggpack_look(oggpack_buffer *b, int bits);
void  oggpack_adv(oggpack_buffer *b, int bits);
long  oggpack_read(oggpack_buffer *b, int bits);
long  oggpack_bytes(oggpack_buffer *b);
unsigned char *oggpack_get_buffer(oggpack_buffer *b);

#endif
