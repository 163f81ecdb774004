/* include/pnmio.h -- PGM/PPM file I/O, same six entry points as the reference's
 * pnmio.h (reference src/V4/pnmio.h:13-49).  Passing img == NULL to the readers
 * makes them malloc the pixel buffer (caller frees). */
#ifndef _PNMIO_H_
#define _PNMIO_H_

#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

unsigned char *pgmReadFile(char *fname, unsigned char *img, int *ncols, int *nrows);
void pgmWriteFile(char *fname, unsigned char *img, int ncols, int nrows);
void ppmWriteFileRGB(char *fname, unsigned char *redimg, unsigned char *greenimg,
                     unsigned char *blueimg, int ncols, int nrows);

unsigned char *pgmRead(FILE *fp, unsigned char *img, int *ncols, int *nrows);
void pgmWrite(FILE *fp, unsigned char *img, int ncols, int nrows);
void ppmWrite(FILE *fp, unsigned char *redimg, unsigned char *greenimg,
              unsigned char *blueimg, int ncols, int nrows);

#ifdef __cplusplus
}
#endif

#endif
