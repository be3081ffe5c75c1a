#include <ctype.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include "mcq_io.h"
#include "mcq_json.h"

/* ---- a small JSON reader: enough for numbers, strings, arrays and objects --------------------- */
typedef enum { J_NULL, J_BOOL, J_NUM, J_STR, J_ARR, J_OBJ } JType;
typedef struct JNode {
  JType type;
  double num;
  char *str;               /* J_STR value, or the member name when the node sits in an object */
  char *key;
  struct JNode *child, *next;
} JNode;

typedef struct { const char *p; } JIn;

static void skip_ws(JIn *in) { while(isspace((unsigned char)*in->p)) in->p++; }

static char *parse_string(JIn *in)
{
  if(*in->p != '"') mcq_die("json: expected a string");
  const char *start = ++in->p;
  while(*in->p && *in->p != '"') { if(*in->p == '\\' && in->p[1]) in->p++; in->p++; }
  if(*in->p != '"') mcq_die("json: unterminated string");
  size_t n = (size_t)(in->p - start);
  char *s = malloc(n + 1);
  memcpy(s, start, n);
  s[n] = '\0';
  in->p++;
  return s;
}

/* Numbers are converted the way the reference's JSON library does it (cJSON.c:96-115: the digits accumulated in a
 * double, times pow(10, exponent)), not with strtod: the two differ in the last bit for most decimal fractions,
 * and a chain resumed from a state file (-r) or run with a parameter file (-p) must start from the very doubles
 * the reference would read. */
static double parse_number(JIn *in)
{
  const char *p = in->p;
  double mant = 0, sign = 1;
  int frac_digits = 0, expo = 0, expo_sign = 1;
  if(*p == '-') { sign = -1; p++; }
  while(isdigit((unsigned char)*p)) mant = mant * 10.0 + (*p++ - '0');
  if(*p == '.' && isdigit((unsigned char)p[1])) {
    p++;
    while(isdigit((unsigned char)*p)) { mant = mant * 10.0 + (*p++ - '0'); frac_digits++; }
  }
  if(*p == 'e' || *p == 'E') {
    p++;
    if(*p == '+') p++;
    else if(*p == '-') { expo_sign = -1; p++; }
    while(isdigit((unsigned char)*p)) expo = expo * 10 + (*p++ - '0');
  }
  in->p = p;
  return sign * mant * pow(10.0, (double)(-frac_digits + expo * expo_sign));
}

static JNode *parse_value(JIn *in)
{
  JNode *n = calloc(1, sizeof(JNode));
  skip_ws(in);
  char c = *in->p;
  if(c == '{' || c == '[') {
    const char close = (c == '{') ? '}' : ']';
    n->type = (c == '{') ? J_OBJ : J_ARR;
    in->p++;
    JNode **tail = &n->child;
    skip_ws(in);
    if(*in->p == close) { in->p++; return n; }
    for(;;) {
      char *key = NULL;
      skip_ws(in);
      if(n->type == J_OBJ) {
        key = parse_string(in);
        skip_ws(in);
        if(*in->p != ':') mcq_die("json: expected ':'");
        in->p++;
      }
      JNode *v = parse_value(in);
      v->key = key;
      *tail = v;
      tail = &v->next;
      skip_ws(in);
      if(*in->p == ',') { in->p++; continue; }
      if(*in->p == close) { in->p++; break; }
      mcq_die("json: expected ',' or '%c'", close);
    }
  } else if(c == '"') {
    n->type = J_STR;
    n->str = parse_string(in);
  } else if(c == '-' || isdigit((unsigned char)c)) {
    n->type = J_NUM;
    n->num = parse_number(in);
  } else if(!strncmp(in->p, "true", 4)) { n->type = J_BOOL; n->num = 1; in->p += 4; }
  else if(!strncmp(in->p, "false", 5)) { n->type = J_BOOL; in->p += 5; }
  else if(!strncmp(in->p, "null", 4)) { n->type = J_NULL; in->p += 4; }
  else mcq_die("json: unexpected character '%c'", c);
  return n;
}

static void jfree(JNode *n)
{
  while(n) {
    JNode *next = n->next;
    jfree(n->child);
    free(n->str); free(n->key); free(n);
    n = next;
  }
}

static JNode *member(const JNode *obj, const char *key, JType type, const char *what)
{
  if(obj == NULL || obj->type != J_OBJ) mcq_die("json: %s: not an object", what);
  for(JNode *c = obj->child; c; c = c->next)
    if(c->key && !strcmp(c->key, key)) {
      if(c->type != type) mcq_die("json: No valid '%s' (%s).", key, what);
      return c;
    }
  mcq_die("json: No '%s' object (%s).", key, what);
}

static char *slurp(const char *path)
{
  gzFile gz = gzopen(path, "r");          /* reads plain and gzip files alike, as the reference does */
  if(gz == NULL) mcq_die("Cannot open file %s.", path);
  size_t cap = 1 << 16, len = 0;
  char *buf = malloc(cap);
  int n;
  while((n = gzread(gz, buf + len, (unsigned)(cap - len - 1))) > 0) {
    len += (size_t)n;
    if(cap - len < 4096) { cap *= 2; buf = realloc(buf, cap); }
  }
  gzclose(gz);
  buf[len] = '\0';
  if(len == 0) mcq_die("Empty json file.");
  return buf;
}

/* load_windows, load_from_json.c:19-45: windows are taken in file order, entry i is swapped with the
 * pool entry of its start site, then the set is sorted and the lengths recomputed */
static void load_windows(const JNode *arr, McqWindowSet *ws, int num_sites)
{
  int count = 0;
  for(const JNode *w = arr->child; w; w = w->next) count++;
  if(count == 0) mcq_die("json: empty window list");
  int *start = malloc(sizeof(int) * count);
  double *coeff = malloc(sizeof(double) * count);
  int i = 0;
  for(const JNode *w = arr->child; w; w = w->next, i++) {
    start[i] = (int)member(w, "start", J_NUM, "window")->num;
    coeff[i] = member(w, "coeff", J_NUM, "window")->num;
    if(start[i] >= num_sites || start[i] < 0) mcq_die("Windows greater than num sites.");
  }
  if(start[0] != 0) mcq_die("json: the first window must start at site 0");
  mcq_windows_set(ws, num_sites, count, start, coeff);
  free(start); free(coeff);
}

void mcq_load_state_json(const char *path, double *llk, double *mu, double *R, double *omega,
                         McqHlaWindows *hw, int num_sites, int num_hla, int num_query, int closest_n,
                         int *closest_refs)
{
  char *text = slurp(path);
  JIn in = {text};
  JNode *root = parse_value(&in);
  const JNode *chain = member(root, "chain", J_ARR, "root");
  const JNode *state = chain->child;
  if(state == NULL) mcq_die("No states in chain.");
  int num_states = 1;
  while(state->next) { state = state->next; num_states++; }
  printf("We got %i states - loading last one.\n", num_states);
  printf("Got state number: %i\n", (int)member(state, "state", J_NUM, "state")->num);
  *mu = member(state, "mu", J_NUM, "state")->num;
  *llk = member(state, "llkhood", J_NUM, "state")->num;

  int s = 0;
  for(const JNode *site = member(state, "sites", J_ARR, "state")->child; site; site = site->next, s++) {
    if(s >= num_sites) mcq_die("Mismatch in sites: more than %i.", num_sites);
    omega[s] = member(site, "omega", J_NUM, "site")->num;
    R[s] = member(site, "R", J_NUM, "site")->num;
  }
  if(s != num_sites) mcq_die("Mismatch in sites: [%i vs %i].", num_sites, s);

  int h = 0;
  for(const JNode *hla = member(state, "hlas", J_ARR, "state")->child; hla; hla = hla->next, h++) {
    if(h >= num_hla) mcq_die("Mismatch in HLAs: more than %i.", num_hla);
    load_windows(member(hla, "esc", J_ARR, "hla"), hw[h].esc, num_sites);
    load_windows(member(hla, "rev", J_ARR, "hla"), hw[h].rev, num_sites);
  }
  if(h != num_hla) mcq_die("Mismatch in HLAs: [%i vs %i].", num_hla, h);

  int q = 0;
  for(const JNode *qr = member(root, "query", J_ARR, "root")->child; qr; qr = qr->next, q++) {
    if(q >= num_query) mcq_die("Mismatch in number of queries: more than %i.", num_query);
    int i = 0;
    for(const JNode *e = member(qr, "closest_n", J_ARR, "query")->child; e; e = e->next, i++) {
      if(i >= closest_n || e->type != J_NUM) mcq_die("Mismatch in closest N.");
      closest_refs[(size_t)q * closest_n + i] = (int)e->num;
    }
    if(i != closest_n) mcq_die("Mismatch in closest N.");
  }
  if(q != num_query) mcq_die("Mismatch in number of queries: [%i vs %i].", q, num_query);
  jfree(root);
  free(text);
}

/* load_fixed_parameters_from_json, load_from_json.c:248-295: kappa, phi and the 64 codon priors */
void mcq_load_fixed_parameters_json(const char *path, double *kappa, double *phi, double *codon_prior)
{
  char *text = slurp(path);
  JIn in = {text};
  JNode *root = parse_value(&in);
  *kappa = member(root, "kappa", J_NUM, "parameters")->num;
  *phi = member(root, "phi", J_NUM, "parameters")->num;
  const JNode *codons = member(root, "codons", J_ARR, "parameters");
  if(codons->child == NULL) mcq_die("json: No 'codons' object.");
  int c = 0;
  for(const JNode *cd = codons->child; cd; cd = cd->next, c++) {
    if(c >= 64) mcq_die("Mismatch in number of codon types: more than 64.");
    codon_prior[c] = member(cd, "codon_prior", J_NUM, "codon")->num;
    printf("codon_prior[%i]:%f\n", c, codon_prior[c]);
  }
  if(c != 64) mcq_die("Mismatch in number of codon types: [%i vs %i].", 64, c);
  jfree(root);
  free(text);
}

/* load_closest_n_from_json, load_from_json.c:178-246: per simulated query the references it was copied from */
void mcq_load_true_mosaics_json(const char *path, int num_query, int num_sites, int *which_seqs, int *num_seqs)
{
  char *text = slurp(path);
  JIn in = {text};
  JNode *root = parse_value(&in);
  const JNode *sims = member(root, "true_sequences", J_ARR, "root");
  if(sims->child == NULL) mcq_die("json: No sequences in file.");
  printf("num sims: %i\n", num_query);
  int q = 0;
  for(const JNode *sim = sims->child; sim; sim = sim->next, q++) {
    if(q >= num_query) mcq_die("json: more mosaics than query sequences [%i].", num_query);
    (void)member(sim, "number", J_NUM, "mosaic");
    int n = 0;
    for(const JNode *e = member(sim, "closest_n", J_ARR, "mosaic")->child; e; e = e->next, n++) {
      if(n >= num_sites) mcq_die("json: num_seqs too big.");
      if(e->type != J_NUM) mcq_die("json: No 'num_seqs' object.");
      which_seqs[(size_t)q * num_sites + n] = (int)e->num;
    }
    num_seqs[q] = n;
  }
  if(q != num_query) mcq_die("json: %i mosaics for %i query sequences.", q, num_query);
  printf("We got %i simulated sequence mosaics\n", q - 1);
  jfree(root);
  free(text);
}
