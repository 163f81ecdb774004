/* klt_error.c -- error convention of the KLT API.
 * Behaviour follows reference src/V1/error.c:23-55: KLTError prints
 * "KLT Error: <msg>\n" to stderr and exit(1)s, KLTWarning prints
 * "KLT Warning: <msg>\n" and returns. */
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include "klt.h"

static void emit(const char *tag, const char *fmt, va_list ap)
{
  fputs(tag, stderr);
  vfprintf(stderr, fmt, ap);
  fputc('\n', stderr);
  fflush(stderr);
}

void KLTError(char *fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  emit("KLT Error: ", fmt, ap);
  va_end(ap);
  exit(1);
}

void KLTWarning(char *fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  emit("KLT Warning: ", fmt, ap);
  va_end(ap);
}
