/* klt_write.c -- feature list / history / table files and the PPM overlay.
 *
 * File formats are those of reference src/V1/writeFeatures.c, byte for byte:
 *  - overlay: 3x3 red squares at (int)(x+0.5),(int)(y+0.5) on the grey image
 *    (:36-89);
 *  - text: the banner, "KLT Feature List|History|Table", counts, a ruler whose
 *    width is the expanded width of "(fmt,fmt)=%5d " (:92-279, :326-401);
 *  - binary: "KLTFL1"/"KLTFH1"/"KLTFT1", int count(s), then per feature
 *    {float x, float y, int val} (:294-301, :340-345, :431-441);
 *  - readers accept both (:446-743).
 * tests/test_host.py checks the writers against the reference's golden
 * features2.txt / features2.ft.
 */
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "klt_internal.h"
#include "pnmio.h"

enum kind { K_LIST, K_HISTORY, K_TABLE };

static const char *const k_title[3] = { "KLT Feature List", "KLT Feature History", "KLT Feature Table" };
static const char *const k_magic[3] = { "KLTFL1", "KLTFH1", "KLTFT1" };
static const char k_bang[] = "!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!\n";
static const char k_warning[] = "!!! Warning:  This is a KLT data file.  "
                                "Do not modify below this line !!!\n";

/* ---- overlay -------------------------------------------------------------------- */
void KLTWriteFeatureListToPPM(KLT_FeatureList fl, KLT_PixelType *greyimg, int ncols, int nrows,
                              char *filename)
{
  const size_t npix = (size_t)ncols * (size_t)nrows;
  unsigned char *r, *g, *b;
  int i, xx, yy;

  if (KLT_verbose >= 1)
    fprintf(stderr, "(KLT) Writing %d features to PPM file: '%s'\n",
            KLTCountRemainingFeatures(fl), filename);
  r = (unsigned char *)malloc(npix ? npix : 1);
  g = (unsigned char *)malloc(npix ? npix : 1);
  b = (unsigned char *)malloc(npix ? npix : 1);
  if (!r || !g || !b) KLTError("(KLTWriteFeaturesToPPM)  Out of memory\n");
  memcpy(r, greyimg, npix);
  memcpy(g, greyimg, npix);
  memcpy(b, greyimg, npix);
  for (i = 0; i < fl->nFeatures; i++) {
    int cx, cy;
    if (fl->feature[i]->val < 0) continue;
    cx = (int)(fl->feature[i]->x + 0.5);
    cy = (int)(fl->feature[i]->y + 0.5);
    for (yy = cy - 1; yy <= cy + 1; yy++)
      for (xx = cx - 1; xx <= cx + 1; xx++)
        if (xx >= 0 && yy >= 0 && xx < ncols && yy < nrows) {
          const size_t o = (size_t)yy * ncols + xx;
          r[o] = 255; g[o] = 0; b[o] = 0;
        }
  }
  ppmWriteFileRGB(filename, r, g, b, ncols, nrows);
  free(r); free(g); free(b);
}

/* ---- text helpers ----------------------------------------------------------------- */
/* Width of a printf format once expanded: "%<n>...<conv>" counts n, "%c" counts
 * one, every other character counts one (writeFeatures.c:173-211). */
static int expanded_width(const char *s)
{
  int w = 0;
  size_t i = 0, len = strlen(s);
  while (i < len) {
    if (s[i] != '%') { w++; i++; continue; }
    if (isdigit((unsigned char)s[i + 1])) {
      w += atoi(s + i + 1);
      i += 2;
      while (i < len && strchr("diouxefgn", s[i]) == NULL) i++;
      if (i >= len) KLTError("(_findStringWidth) Can't determine length of string '%s'", s);
      i++;
    } else if (s[i + 1] == 'c') {
      w++; i += 2;
    } else {
      KLTError("(_findStringWidth) Can't determine length of string '%s'", s);
    }
  }
  return w;
}

static void hyphens(FILE *fp, int n) { while (n-- > 0) fputc('-', fp); }

static FILE *open_text(char *fname, char *fmt, char *cell, char *type)
{
  FILE *fp = (fname == NULL) ? stderr : fopen(fname, "wb");
  size_t n;
  if (fp == NULL) KLTError("(KLTWriteFeatures) Can't open file '%s' for writing\n", fname);
  if (fmt[0] != '%') KLTError("(KLTWriteFeatures) Bad Format: %s\n", fmt);
  n = strlen(fmt);
  *type = fmt[n - 1];
  if (*type != 'f' && *type != 'd') KLTError("(KLTWriteFeatures) Format must end in 'f' or 'd'.");
  sprintf(cell, "(%s,%s)=%%%dd ", fmt, fmt, 5);
  return fp;
}

static FILE *open_binary(char *fname)
{
  FILE *fp;
  if (fname == NULL) KLTError("(KLTWriteFeatures) Can't write binary data to stderr");
  fp = fopen(fname, "wb");
  if (fp == NULL) KLTError("(KLTWriteFeatures) Can't open file '%s' for writing", fname);
  return fp;
}

static void text_header(FILE *fp, const char *cell, enum kind k, int nFrames, int nFeatures)
{
  const int w = expanded_width(cell);
  int i;
  if (fp != stderr) {
    fputs("Feel free to place comments here.\n\n\n", fp);
    fputs(k_bang, fp);
    fputs(k_warning, fp);
    fputs("\n", fp);
  }
  fputs("------------------------------\n", fp);
  fprintf(fp, "%s\n", k_title[k]);
  fputs("------------------------------\n\n", fp);
  if (k == K_LIST) fprintf(fp, "nFeatures = %d\n\n", nFeatures);
  else if (k == K_HISTORY) fprintf(fp, "nFrames = %d\n\n", nFrames);
  else fprintf(fp, "nFrames = %d, nFeatures = %d\n\n", nFrames, nFeatures);

  if (k == K_LIST) {
    fputs("feature | (x,y)=val\n--------+-", fp);
    hyphens(fp, w);
    fputc('\n', fp);
  } else if (k == K_HISTORY) {
    fputs("frame | (x,y)=val\n------+-", fp);
    hyphens(fp, w);
    fputc('\n', fp);
  } else {
    fputs("feature |          frame\n        |", fp);
    for (i = 0; i < nFrames; i++) fprintf(fp, "%*d", w, i);
    fputs("\n--------+-", fp);
    for (i = 0; i < nFrames; i++) hyphens(fp, w);
    fputc('\n', fp);
  }
}

static void text_cell(FILE *fp, KLT_Feature f, const char *cell, char type)
{
  if (type == 'f') {
    fprintf(fp, cell, (float)f->x, (float)f->y, f->val);
  } else {
    float x = f->x, y = f->y;           /* round to nearest unless negative */
    if (x >= 0.0) x += 0.5;
    if (y >= 0.0) y += 0.5;
    fprintf(fp, cell, (int)x, (int)y, f->val);
  }
}

static void binary_cell(FILE *fp, KLT_Feature f)
{
  fwrite(&f->x, sizeof(KLT_locType), 1, fp);
  fwrite(&f->y, sizeof(KLT_locType), 1, fp);
  fwrite(&f->val, sizeof(int), 1, fp);
}

static void announce(const char *what, char *fname, char *fmt)
{
  if (KLT_verbose >= 1 && fname != NULL)
    fprintf(stderr, "(KLT) Writing feature %s to %s file: '%s'\n", what,
            fmt == NULL ? "binary" : "text", fname);
}

/* ---- writers ------------------------------------------------------------------------ */
void KLTWriteFeatureList(KLT_FeatureList fl, char *fname, char *fmt)
{
  FILE *fp;
  int i;
  announce("list", fname, fmt);
  if (fmt != NULL) {
    char cell[100], type;
    fp = open_text(fname, fmt, cell, &type);
    text_header(fp, cell, K_LIST, 0, fl->nFeatures);
    for (i = 0; i < fl->nFeatures; i++) {
      fprintf(fp, "%7d | ", i);
      text_cell(fp, fl->feature[i], cell, type);
      fputc('\n', fp);
    }
    if (fp != stderr) fclose(fp);
  } else {
    fp = open_binary(fname);
    fwrite(k_magic[K_LIST], 1, 6, fp);
    fwrite(&fl->nFeatures, sizeof(int), 1, fp);
    for (i = 0; i < fl->nFeatures; i++) binary_cell(fp, fl->feature[i]);
    fclose(fp);
  }
}

void KLTWriteFeatureHistory(KLT_FeatureHistory fh, char *fname, char *fmt)
{
  FILE *fp;
  int i;
  announce("history", fname, fmt);
  if (fmt != NULL) {
    char cell[100], type;
    fp = open_text(fname, fmt, cell, &type);
    text_header(fp, cell, K_HISTORY, fh->nFrames, 0);
    for (i = 0; i < fh->nFrames; i++) {
      fprintf(fp, "%5d | ", i);
      text_cell(fp, fh->feature[i], cell, type);
      fputc('\n', fp);
    }
    if (fp != stderr) fclose(fp);
  } else {
    fp = open_binary(fname);
    fwrite(k_magic[K_HISTORY], 1, 6, fp);
    fwrite(&fh->nFrames, sizeof(int), 1, fp);
    for (i = 0; i < fh->nFrames; i++) binary_cell(fp, fh->feature[i]);
    fclose(fp);
  }
}

void KLTWriteFeatureTable(KLT_FeatureTable ft, char *fname, char *fmt)
{
  FILE *fp;
  int i, j;
  announce("table", fname, fmt);
  if (fmt != NULL) {
    char cell[100], type;
    fp = open_text(fname, fmt, cell, &type);
    text_header(fp, cell, K_TABLE, ft->nFrames, ft->nFeatures);
    for (j = 0; j < ft->nFeatures; j++) {
      fprintf(fp, "%7d | ", j);
      for (i = 0; i < ft->nFrames; i++) text_cell(fp, ft->feature[j][i], cell, type);
      fputc('\n', fp);
    }
    if (fp != stderr) fclose(fp);
  } else {
    fp = open_binary(fname);
    fwrite(k_magic[K_TABLE], 1, 6, fp);
    fwrite(&ft->nFrames, sizeof(int), 1, fp);
    fwrite(&ft->nFeatures, sizeof(int), 1, fp);
    for (j = 0; j < ft->nFeatures; j++)
      for (i = 0; i < ft->nFrames; i++) binary_cell(fp, ft->feature[j][i]);
    fclose(fp);
  }
}

/* ---- readers ------------------------------------------------------------------------ */
static void skip_past(FILE *fp, int ch)
{
  int c;
  do { c = fgetc(fp); } while (c != ch && c != EOF);
}

static void expect_word(FILE *fp, const char *want)
{
  char w[100];
  if (fscanf(fp, "%99s", w) != 1 || strcmp(w, want) != 0)
    KLTError("(_readFeatures) File is corrupted -- (Expected '%s', found '%s' instead)", want, w);
}

/* Returns the kind of file; fills whichever of nFrames/nFeatures is non-NULL. */
static enum kind read_header(FILE *fp, int *nFrames, int *nFeatures, int *binary)
{
  char line[100];
  enum kind k;
  size_t got = fread(line, 1, 6, fp);
  line[got] = '\0';
  *binary = 1;
  if (strcmp(line, k_magic[K_LIST]) == 0) {
    if (nFeatures && fread(nFeatures, sizeof(int), 1, fp) != 1) *nFeatures = 0;
    return K_LIST;
  }
  if (strcmp(line, k_magic[K_HISTORY]) == 0) {
    if (nFrames && fread(nFrames, sizeof(int), 1, fp) != 1) *nFrames = 0;
    return K_HISTORY;
  }
  if (strcmp(line, k_magic[K_TABLE]) == 0) {
    if (nFrames && fread(nFrames, sizeof(int), 1, fp) != 1) *nFrames = 0;
    if (nFeatures && fread(nFeatures, sizeof(int), 1, fp) != 1) *nFeatures = 0;
    return K_TABLE;
  }
  *binary = 0;
  rewind(fp);
  do {
    if (fgets(line, sizeof line, fp) == NULL)
      KLTError("(_readFeatures) File is corrupted -- Couldn't find line:\n\t%s\n", k_warning);
  } while (strcmp(line, k_warning) != 0);
  skip_past(fp, '-');
  skip_past(fp, '\n');
  if (fgets(line, sizeof line, fp) == NULL) line[0] = '\0';
  if (strcmp(line, "KLT Feature List\n") == 0) k = K_LIST;
  else if (strcmp(line, "KLT Feature History\n") == 0) k = K_HISTORY;
  else if (strcmp(line, "KLT Feature Table\n") == 0) k = K_TABLE;
  else {
    KLTError("(_readFeatures) File is corrupted -- (Not 'KLT Feature List', "
             "'KLT Feature History', or 'KLT Feature Table')");
    return K_LIST;
  }
  /* wrong container passed: let the caller report it */
  if ((k == K_LIST && !nFeatures) || (k == K_HISTORY && !nFrames) ||
      (k == K_TABLE && (!nFeatures || !nFrames)))
    return k;
  skip_past(fp, '-');
  skip_past(fp, '\n');
  if (k == K_LIST) {
    expect_word(fp, "nFeatures"); expect_word(fp, "=");
    if (fscanf(fp, "%d", nFeatures) != 1) *nFeatures = 0;
  } else {
    expect_word(fp, "nFrames"); expect_word(fp, "=");
    if (fscanf(fp, "%d", nFrames) != 1) *nFrames = 0;
    if (k == K_TABLE) {
      expect_word(fp, ","); expect_word(fp, "nFeatures"); expect_word(fp, "=");
      if (fscanf(fp, "%d", nFeatures) != 1) *nFeatures = 0;
    }
  }
  skip_past(fp, '-');
  skip_past(fp, '\n');
  return k;
}

static void read_text_cell(FILE *fp, KLT_Feature f)
{
  skip_past(fp, '(');
  if (fscanf(fp, "%f,%f)=%d", &f->x, &f->y, &f->val) != 3)
    KLTError("(_readFeatureTxt) File is corrupted -- bad feature entry");
}

static void read_binary_cell(FILE *fp, KLT_Feature f)
{
  if (fread(&f->x, sizeof(KLT_locType), 1, fp) != 1 ||
      fread(&f->y, sizeof(KLT_locType), 1, fp) != 1 ||
      fread(&f->val, sizeof(int), 1, fp) != 1)
    KLTError("(_readFeatureBin) File is truncated");
}

static FILE *open_for_read(const char *who, char *fname, const char *what)
{
  FILE *fp = fopen(fname, "rb");
  if (fp == NULL) KLTError("(%s) Can't open file '%s' for reading", who, fname);
  if (KLT_verbose >= 1) fprintf(stderr, "(KLT) Reading feature %s from '%s'\n", what, fname);
  return fp;
}

static void read_row_index(FILE *fp, const char *who, int want)
{
  int idx = -1;
  if (fscanf(fp, "%d |", &idx) != 1 || idx != want)
    KLTError("(%s) Bad index at i = %d-- %d", who, want, idx);
}

KLT_FeatureList KLTReadFeatureList(KLT_FeatureList fl_in, char *fname)
{
  FILE *fp = open_for_read("KLTReadFeatureList", fname, "list");
  KLT_FeatureList fl = fl_in;
  int n = 0, binary, i;
  if (read_header(fp, NULL, &n, &binary) != K_LIST)
    KLTError("(KLTReadFeatureList) File '%s' does not contain a FeatureList", fname);
  if (fl == NULL) fl = KLTCreateFeatureList(n);
  else if (fl->nFeatures != n)
    KLTError("(KLTReadFeatureList) The feature list passed does not contain the same "
             "number of features as the feature list in file '%s' ", fname);
  for (i = 0; i < fl->nFeatures; i++) {
    if (binary) read_binary_cell(fp, fl->feature[i]);
    else { read_row_index(fp, "KLTReadFeatureList", i); read_text_cell(fp, fl->feature[i]); }
  }
  fclose(fp);
  return fl;
}

KLT_FeatureHistory KLTReadFeatureHistory(KLT_FeatureHistory fh_in, char *fname)
{
  FILE *fp = open_for_read("KLTReadFeatureHistory", fname, "history");
  KLT_FeatureHistory fh = fh_in;
  int n = 0, binary, i;
  if (read_header(fp, &n, NULL, &binary) != K_HISTORY)
    KLTError("(KLTReadFeatureHistory) File '%s' does not contain a FeatureHistory", fname);
  if (fh == NULL) fh = KLTCreateFeatureHistory(n);
  else if (fh->nFrames != n)
    KLTError("(KLTReadFeatureHistory) The feature history passed does not contain the same "
             "number of frames as the feature history in file '%s' ", fname);
  for (i = 0; i < fh->nFrames; i++) {
    if (binary) read_binary_cell(fp, fh->feature[i]);
    else { read_row_index(fp, "KLTReadFeatureHistory", i); read_text_cell(fp, fh->feature[i]); }
  }
  fclose(fp);
  return fh;
}

KLT_FeatureTable KLTReadFeatureTable(KLT_FeatureTable ft_in, char *fname)
{
  FILE *fp = open_for_read("KLTReadFeatureTable", fname, "table");
  KLT_FeatureTable ft = ft_in;
  int nfr = 0, nfe = 0, binary, i, j;
  if (read_header(fp, &nfr, &nfe, &binary) != K_TABLE)
    KLTError("(KLTReadFeatureTable) File '%s' does not contain a FeatureTable", fname);
  if (ft == NULL) ft = KLTCreateFeatureTable(nfr, nfe);
  else if (ft->nFrames != nfr || ft->nFeatures != nfe)
    KLTError("(KLTReadFeatureTable) The feature table passed does not contain the same number "
             "of frames and features as the feature table in file '%s' ", fname);
  for (j = 0; j < ft->nFeatures; j++) {
    if (!binary) read_row_index(fp, "KLTReadFeatureTable", j);
    for (i = 0; i < ft->nFrames; i++) {
      if (binary) read_binary_cell(fp, ft->feature[j][i]);
      else read_text_cell(fp, ft->feature[j][i]);
    }
  }
  fclose(fp);
  return ft;
}
