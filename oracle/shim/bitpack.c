/* LSb-first bit packer with the oggpack_* call surface libvorbis uses.
 * TEST INFRASTRUCTURE ONLY (lets oracle/_ref link without libogg).
 * Semantics follow the Vorbis I spec, doc/02-bitpacking.tex: values of up to
 * 32 bits are appended least-significant-bit first; reads past the end
 * return -1.
 */
#include <string.h>
#include "ogg/ogg.h"

#define CHUNK 256

static const unsigned long lowmask[33] = {
  0x0,0x1,0x3,0x7,0xf,0x1f,0x3f,0x7f,0xff,0x1ff,0x3ff,0x7ff,0xfff,0x1fff,
  0x3fff,0x7fff,0xffff,0x1ffff,0x3ffff,0x7ffff,0xfffff,0x1fffff,0x3fffff,
  0x7fffff,0xffffff,0x1ffffff,0x3ffffff,0x7ffffff,0xfffffff,0x1fffffff,
  0x3fffffff,0x7fffffff,0xffffffff };

void oggpack_writeinit(oggpack_buffer *b){
  memset(b,0,sizeof(*b));
  b->buffer = b->ptr = (unsigned char*)calloc(CHUNK,1);
  b->storage = CHUNK;
}

void oggpack_reset(oggpack_buffer *b){
  if(!b->buffer) return;
  b->ptr = b->buffer;
  b->buffer[0] = 0;
  b->endbit = 0;
  b->endbyte = 0;
}

void oggpack_writeclear(oggpack_buffer *b){
  if(b->buffer) free(b->buffer);
  memset(b,0,sizeof(*b));
}

void oggpack_writetrunc(oggpack_buffer *b, long bits){
  long bytes = bits >> 3;
  if(!b->buffer) return;
  bits &= 7;
  b->ptr = b->buffer + bytes;
  b->endbit = (int)bits;
  b->endbyte = bytes;
  *b->ptr &= (unsigned char)lowmask[bits];
}

void oggpack_write(oggpack_buffer *b, unsigned long value, int bits){
  if(bits < 0 || bits > 32 || !b->buffer) return;
  if(b->endbyte + 8 >= b->storage){
    long ns = b->storage + CHUNK;
    unsigned char *nb = (unsigned char*)realloc(b->buffer, (size_t)ns);
    if(!nb) return;
    memset(nb + b->storage, 0, CHUNK);
    b->buffer = nb;
    b->storage = ns;
    b->ptr = nb + b->endbyte;
  }
  value &= lowmask[bits];
  {
    int have = b->endbit;          /* bits already used in *ptr */
    int total = have + bits;
    unsigned long long acc = ((unsigned long long)value) << have;
    int k = 0;
    b->ptr[0] |= (unsigned char)(acc & 0xff);
    for(k = 1; k*8 < total; k++)
      b->ptr[k] = (unsigned char)((acc >> (8*k)) & 0xff);
    b->endbyte += total / 8;
    b->ptr     += total / 8;
    b->endbit   = total & 7;
    if(total >= 8 && b->endbit == 0) b->ptr[0] = 0;
  }
}

void oggpack_readinit(oggpack_buffer *b, unsigned char *buf, int bytes){
  memset(b,0,sizeof(*b));
  b->buffer = b->ptr = buf;
  b->storage = bytes;
}

long oggpack_look(oggpack_buffer *b, int bits){
  unsigned long long acc = 0;
  int need, k;
  if(bits < 0 || bits > 32) return -1;
  if(b->endbyte*8 + b->endbit + bits > b->storage*8) return -1;
  if(!bits) return 0;
  need = (b->endbit + bits + 7) / 8;
  for(k = 0; k < need; k++)
    acc |= ((unsigned long long)b->ptr[k]) << (8*k);
  return (long)((acc >> b->endbit) & lowmask[bits]);
}

void oggpack_adv(oggpack_buffer *b, int bits){
  long pos = b->endbyte*8 + b->endbit + bits;
  if(pos > b->storage*8){
    b->ptr = NULL; b->endbyte = b->storage; b->endbit = 1; /* overflow */
    return;
  }
  b->ptr = b->buffer + (pos >> 3);
  b->endbyte = pos >> 3;
  b->endbit = (int)(pos & 7);
}

long oggpack_read(oggpack_buffer *b, int bits){
  long v;
  if(!b->ptr) return -1;
  v = oggpack_look(b, bits);
  if(v < 0 && bits){ b->ptr = NULL; b->endbyte = b->storage; b->endbit = 1; return -1; }
  oggpack_adv(b, bits);
  return v;
}

long oggpack_bytes(oggpack_buffer *b){
  return b->endbyte + (b->endbit + 7) / 8;
}

unsigned char *oggpack_get_buffer(oggpack_buffer *b){
  return b->buffer;
}
