/* TEST INFRASTRUCTURE ONLY (oracle build): header-only stand-in for the
 * un-vendored noporpoise/seq_file library (libs/Makefile:22-23 of the
 * reference).  FASTA only ('>' header lines, sequence possibly wrapped over
 * several lines), which is all load_data.c:66-94 needs here.
 */
#ifndef MCQ_SHIM_SEQ_FILE_H_
#define MCQ_SHIM_SEQ_FILE_H_

#include "string_buffer.h"

typedef struct { StrBuf name, seq, qual; } read_t;
typedef struct { FILE *fh; int pending; /* next unread char or -2 */ } seq_file_t;

static inline read_t *seq_read_alloc(read_t *r)
{
  strbuf_alloc(&r->name, 256);
  strbuf_alloc(&r->seq, 4096);
  strbuf_alloc(&r->qual, 16);
  return r;
}

static inline void seq_read_dealloc(read_t *r)
{
  strbuf_dealloc(&r->name); strbuf_dealloc(&r->seq); strbuf_dealloc(&r->qual);
}

static inline seq_file_t *seq_open(const char *path)
{
  FILE *fh = fopen(path, "r");
  if(fh == NULL) return NULL;
  seq_file_t *sf = malloc(sizeof(seq_file_t));
  sf->fh = fh;
  return sf;
}

static inline void seq_close(seq_file_t *sf) { fclose(sf->fh); free(sf); }

/* Returns 1 if a record was read, 0 at end of file. */
static inline int seq_read(seq_file_t *sf, read_t *r)
{
  int c;
  strbuf_reset(&r->name); strbuf_reset(&r->seq);
  /* find header */
  while((c = fgetc(sf->fh)) != EOF && c != '>') {}
  if(c == EOF) return 0;
  strbuf_readline(&r->name, sf->fh);
  strbuf_chomp(&r->name);
  /* sequence lines until next '>' or EOF */
  while((c = fgetc(sf->fh)) != EOF) {
    if(c == '>') { ungetc(c, sf->fh); break; }
    if(c == '\n' || c == '\r' || c == ' ' || c == '\t') continue;
    strbuf_ensure_(&r->seq, 1);
    r->seq.b[r->seq.end++] = (char)c;
  }
  r->seq.b[r->seq.end] = '\0';
  return 1;
}

#endif
