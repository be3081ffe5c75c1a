/* TEST INFRASTRUCTURE ONLY (oracle build): header-only stand-in for the
 * un-vendored noporpoise/string_buffer library (libs/Makefile:25-26 of the
 * reference clones it at an unpinned HEAD).  Only what load_data.c and
 * load_from_json.c use.  Parsing only, no arithmetic.
 */
#ifndef MCQ_SHIM_STRING_BUFFER_H_
#define MCQ_SHIM_STRING_BUFFER_H_

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include <zlib.h>

typedef struct { char *b; size_t end, size; } StrBuf;

static inline StrBuf *strbuf_alloc(StrBuf *sb, size_t cap)
{
  if(cap < 16) cap = 16;
  sb->b = malloc(cap);
  sb->size = cap;
  sb->end = 0;
  sb->b[0] = '\0';
  return sb;
}

static inline void strbuf_dealloc(StrBuf *sb) { free(sb->b); sb->b = NULL; sb->end = sb->size = 0; }

static inline void strbuf_reset(StrBuf *sb) { sb->end = 0; sb->b[0] = '\0'; }

static inline void strbuf_ensure_(StrBuf *sb, size_t extra)
{
  if(sb->end + extra + 1 > sb->size) {
    while(sb->end + extra + 1 > sb->size) sb->size *= 2;
    sb->b = realloc(sb->b, sb->size);
  }
}

/* Append one line (including its newline) from fh; returns chars read. */
static inline size_t strbuf_readline(StrBuf *sb, FILE *fh)
{
  size_t start = sb->end;
  int c;
  while((c = fgetc(fh)) != EOF) {
    strbuf_ensure_(sb, 1);
    sb->b[sb->end++] = (char)c;
    if(c == '\n') break;
  }
  sb->b[sb->end] = '\0';
  return sb->end - start;
}

static inline size_t strbuf_reset_readline(StrBuf *sb, FILE *fh)
{
  strbuf_reset(sb);
  return strbuf_readline(sb, fh);
}

static inline size_t strbuf_gzreadline(StrBuf *sb, gzFile gz)
{
  size_t start = sb->end;
  int c;
  while((c = gzgetc(gz)) != -1) {
    strbuf_ensure_(sb, 1);
    sb->b[sb->end++] = (char)c;
    if(c == '\n') break;
  }
  sb->b[sb->end] = '\0';
  return sb->end - start;
}

static inline void strbuf_chomp(StrBuf *sb)
{
  while(sb->end > 0 && (sb->b[sb->end-1] == '\n' || sb->b[sb->end-1] == '\r')) sb->end--;
  sb->b[sb->end] = '\0';
}

#endif
