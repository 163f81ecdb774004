/* pnmio.c -- binary PGM (P5) / PPM (P6) readers and writers.
 * Same entry points and file format handling as reference src/V1/pnmio.c:
 * header tokens may be separated by any whitespace and '#' comments
 * (:20-77), dimensions above 10000 are refused (:66), maxval other than 255
 * only warns (:75), img == NULL makes the reader allocate (:157-166).
 * Writers emit "P5\n<w> <h>\n255\n" / "P6\n..." and raw rows; the PPM writer
 * interleaves through a row buffer instead of three fwrite calls per pixel
 * (same bytes on disk). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "klt.h"
#include "pnmio.h"

/* next header token, skipping whitespace and # comments; "" at EOF */
/* returns 1 when the token was ended by a comment that abuts it ("255#c\n"): the comment's
 * newline, which then is the single separator after the token, has already been consumed */
static int next_token(FILE *fp, char *tok, int cap)
{
  int c, n = 0, ate_separator = 0;
  tok[0] = '\0';
  for (;;) {
    c = fgetc(fp);
    if (c == EOF) break;
    if (c == '#') {
      while (c != '\n' && c != EOF) c = fgetc(fp);
      if (n > 0) { ate_separator = 1; break; }
      continue;
    }
    if (c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v') {
      if (n > 0) { ungetc(c, fp); break; }
      continue;
    }
    if (n < cap - 1) tok[n++] = (char)c;
  }
  tok[n] = '\0';
  return ate_separator;
}

void pnmReadHeader(FILE *fp, int *magic, int *ncols, int *nrows, int *maxval)
{
  char tok[80];
  next_token(fp, tok, sizeof tok);
  if (tok[0] != 'P')
    KLTError("(pnmReadHeader) Magic number does not begin with 'P', but with a '%c'", tok[0]);
  sscanf(tok, "P%d", magic);
  next_token(fp, tok, sizeof tok);
  *ncols = atoi(tok);
  next_token(fp, tok, sizeof tok);
  *nrows = atoi(tok);
  if (*ncols < 0 || *nrows < 0 || *ncols > 10000 || *nrows > 10000)
    KLTError("(pnmReadHeader) The dimensions %d x %d are unacceptable", *ncols, *nrows);
  /* the single whitespace after maxval -- unless a comment abutting maxval already took its
   * newline (the reference, pnmio.c:66-69, reads one byte regardless and loses the first pixel) */
  if (!next_token(fp, tok, sizeof tok)) fgetc(fp);
  *maxval = atoi(tok);
  if (*maxval != 255)
    KLTWarning("(pnmReadHeader) Maxval is not 255, but %d", *maxval);
}

void pgmReadHeader(FILE *fp, int *magic, int *ncols, int *nrows, int *maxval)
{
  pnmReadHeader(fp, magic, ncols, nrows, maxval);
  if (*magic != 5) KLTError("(pgmReadHeader) Magic number is not 'P5', but 'P%d'", *magic);
}

void ppmReadHeader(FILE *fp, int *magic, int *ncols, int *nrows, int *maxval)
{
  pnmReadHeader(fp, magic, ncols, nrows, maxval);
  if (*magic != 6) KLTError("(ppmReadHeader) Magic number is not 'P6', but 'P%d'", *magic);
}

static FILE *open_or_die(const char *who, const char *fname, const char *mode)
{
  FILE *fp = fopen(fname, mode);
  if (fp == NULL)
    KLTError("(%s) Can't open file named '%s' for %s\n", who, fname,
             mode[0] == 'r' ? "reading" : "writing");
  return fp;
}

void pgmReadHeaderFile(char *fname, int *magic, int *ncols, int *nrows, int *maxval)
{
  FILE *fp = open_or_die("pgmReadHeaderFile", fname, "rb");
  pgmReadHeader(fp, magic, ncols, nrows, maxval);
  fclose(fp);
}

void ppmReadHeaderFile(char *fname, int *magic, int *ncols, int *nrows, int *maxval)
{
  FILE *fp = open_or_die("ppmReadHeaderFile", fname, "rb");
  ppmReadHeader(fp, magic, ncols, nrows, maxval);
  fclose(fp);
}

unsigned char *pgmRead(FILE *fp, unsigned char *img, int *ncols, int *nrows)
{
  int magic, maxval;
  size_t npix;
  unsigned char *dst = img;
  pgmReadHeader(fp, &magic, ncols, nrows, &maxval);
  npix = (size_t)*ncols * (size_t)*nrows;
  if (dst == NULL) {
    dst = (unsigned char *)malloc(npix ? npix : 1);
    if (dst == NULL) KLTError("(pgmRead) Memory not allocated");
  }
  if (fread(dst, 1, npix, fp) != npix) { /* short file: keep what was read, as the reference does */ }
  return dst;
}

unsigned char *pgmReadFile(char *fname, unsigned char *img, int *ncols, int *nrows)
{
  FILE *fp = open_or_die("pgmReadFile", fname, "rb");
  unsigned char *p = pgmRead(fp, img, ncols, nrows);
  fclose(fp);
  return p;
}

void pgmWrite(FILE *fp, unsigned char *img, int ncols, int nrows)
{
  fprintf(fp, "P5\n%d %d\n255\n", ncols, nrows);
  fwrite(img, 1, (size_t)ncols * (size_t)nrows, fp);
}

void pgmWriteFile(char *fname, unsigned char *img, int ncols, int nrows)
{
  FILE *fp = open_or_die("pgmWriteFile", fname, "wb");
  pgmWrite(fp, img, ncols, nrows);
  fclose(fp);
}

void ppmWrite(FILE *fp, unsigned char *redimg, unsigned char *greenimg, unsigned char *blueimg,
              int ncols, int nrows)
{
  unsigned char *row = (unsigned char *)malloc(3 * (size_t)(ncols > 0 ? ncols : 1));
  int x, y;
  if (row == NULL) KLTError("(ppmWrite) Out of memory");
  fprintf(fp, "P6\n%d %d\n255\n", ncols, nrows);
  for (y = 0; y < nrows; y++) {
    const size_t o = (size_t)y * ncols;
    for (x = 0; x < ncols; x++) {
      row[3 * x] = redimg[o + x];
      row[3 * x + 1] = greenimg[o + x];
      row[3 * x + 2] = blueimg[o + x];
    }
    fwrite(row, 3, (size_t)ncols, fp);
  }
  free(row);
}

void ppmWriteFileRGB(char *fname, unsigned char *redimg, unsigned char *greenimg,
                     unsigned char *blueimg, int ncols, int nrows)
{
  FILE *fp = open_or_die("ppmWriteFileRGB", fname, "wb");
  ppmWrite(fp, redimg, greenimg, blueimg, ncols, nrows);
  fclose(fp);
}
