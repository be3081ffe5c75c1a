#define _GNU_SOURCE
#include <ctype.h>
#include <errno.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include "mcq_io.h"

void mcq_die(const char *fmt, ...)
{
  va_list ap;
  fflush(stdout);
  fprintf(stderr, "[mcq_dmodel] Error: ");
  va_start(ap, fmt);
  vfprintf(stderr, fmt, ap);
  va_end(ap);
  fputc('\n', stderr);
  exit(EXIT_FAILURE);
}

int mcq_load_fasta(const char *path, char ***out)
{
  FILE *fh = fopen(path, "r");
  if(!fh) mcq_die("Cannot open file: %s.", path);
  int cap = 1024, n = 0;
  char **seqs = malloc(sizeof(char *) * cap);
  char *line = NULL;
  size_t lcap = 0;
  ssize_t len;
  char *cur = NULL;
  size_t clen = 0, ccap = 0;
  int in_record = 0;
  while((len = getline(&line, &lcap, fh)) >= 0) {
    if(len > 0 && line[0] == '>') {
      if(in_record) {
        if(n == cap) { cap *= 2; seqs = realloc(seqs, sizeof(char *) * cap); }
        seqs[n++] = cur ? cur : strdup("");
      }
      cur = NULL; clen = ccap = 0; in_record = 1;
      continue;
    }
    if(!in_record) continue;
    for(ssize_t i = 0; i < len; i++) {
      char c = line[i];
      if(c == '\n' || c == '\r' || c == ' ' || c == '\t') continue;
      if(clen + 2 > ccap) { ccap = ccap ? 2 * ccap : 4096; cur = realloc(cur, ccap); }
      cur[clen++] = c;
      cur[clen] = '\0';
    }
  }
  if(in_record) {
    if(n == cap) { cap *= 2; seqs = realloc(seqs, sizeof(char *) * cap); }
    seqs[n++] = cur ? cur : strdup("");
  }
  free(line);
  fclose(fh);
  *out = seqs;
  return n;
}

int mcq_load_hla_csv(const char *path, char ***out, int num_rows)
{
  FILE *fh = fopen(path, "r");
  if(!fh) mcq_die("Cannot open file: %s.", path);
  char *line = NULL;
  size_t lcap = 0;
  if(getline(&line, &lcap, fh) <= 0) mcq_die("Empty CSV file: %s.", path);
  int types = 0;
  for(char *p = line; *p; p++) types += (*p == ',');
  char **rows = malloc(sizeof(char *) * num_rows);
  char *data = malloc((size_t)num_rows * (types + 1));
  int i;
  for(i = 0; i < num_rows && getline(&line, &lcap, fh) > 0; i++) {
    rows[i] = data + (size_t)i * (types + 1);
    char *p = strchr(line, ',');
    for(int h = 0; h < types; h++) {
      if(p == NULL || *p != ',') mcq_die("Missing CSV entries.");
      p++;
      while(isspace((unsigned char)*p)) p++;
      if(*p != '0' && *p != '1') mcq_die("Invalid CSV line: %s.", p);
      rows[i][h] = *p++;
      while(isspace((unsigned char)*p)) p++;
    }
    if(p != NULL && *p != '\0') mcq_die("Extra CSV columns.");
    rows[i][types] = '\0';
  }
  if(i < num_rows) mcq_die("Not enough rows in CSV file: %s.", path);
  free(line);
  fclose(fh);
  *out = rows;
  return types;
}

int mcq_mkpath(const char *path)
{
  char *copy = strdup(path);
  int ok = 1;
  for(char *p = copy + 1; ; p++) {
    if(*p == '/' || *p == '\0') {
      char keep = *p;
      *p = '\0';
      struct stat st;
      if(stat(copy, &st) != 0) { if(mkdir(copy, 0755) != 0 && errno != EEXIST) { ok = 0; break; } }
      else if(!S_ISDIR(st.st_mode)) { ok = 0; break; }
      *p = keep;
      if(keep == '\0') break;
    }
  }
  free(copy);
  return ok;
}

/* the layout write_json.c:9-67 produces, so that a state written here can be resumed by the reference */
void mcq_write_state_json(FILE *fh, int state, double llk, double mu, int num_sites, const double *omega,
                          const double *R, const McqHlaWindows *hw, int num_hla, int num_query,
                          int closest_n, const int *closest_refs)
{
  fprintf(fh, "{\n \"chain\":[{\n    \"state\":%i,\n    \"llkhood\": %.10f,\n    \"mu\": %.10f,\n    \"sites\": [",
          state, llk, mu);
  for(int i = 0; i < num_sites; i++) {
    if(i != 0) fprintf(fh, "              ");
    fprintf(fh, "{\"omega\":%.10f, \"R\":%.10f}%s\n", omega[i], R[i], i != num_sites - 1 ? "," : "],");
  }
  fprintf(fh, "    \"hlas\": [{\n");
  for(int h = 0; h < num_hla; h++) {
    if(h != 0) fprintf(fh, "    {\n");
    fprintf(fh, "      \"id\": %i,\n      \"esc\": [", h + 1);
    const McqWindowSet *sets[2] = {hw[h].esc, hw[h].rev};
    for(int k = 0; k < 2; k++) {
      const McqWindowSet *ws = sets[k];
      if(k == 1) fprintf(fh, "      \"rev\": [");
      for(int w = 0; w < ws->count; w++) {
        if(w != 0) fprintf(fh, "                      ");
        fprintf(fh, "{\"start\":%i, \"len\":%i, \"coeff\":%.10f}%s\n", ws->w[w].start, ws->w[w].length,
                ws->w[w].coeff, w != ws->count - 1 ? "," : (k == 0 ? "]," : "]"));
      }
    }
    fprintf(fh, "%s\n", h != num_hla - 1 ? "    }," : "    }]");
  }
  fprintf(fh, "  }],\n \"query\":[{\n");
  for(int q = 0; q < num_query; q++) {
    fprintf(fh, "    \"q\":%i,\n    \"closest_n\": [", q);
    for(int n = 0; n < closest_n; n++)
      fprintf(fh, "%i%s", closest_refs[(size_t)q * closest_n + n], n != closest_n - 1 ? "," : "]");
    fprintf(fh, "%s\n", q != num_query - 1 ? "\n  },\n    {" : "\n  }]");
  }
  fprintf(fh, "}");
}
