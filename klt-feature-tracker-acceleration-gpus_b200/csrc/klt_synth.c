/* klt_synth.c -- deterministic synthetic image sequences for benches and tests
 * (BASELINE.json configs 4 and 5; SURVEY.md 8d "Config 4").  NOT on the hot
 * path: plain C on the host, built into its own libklt_synth.so so that both
 * bench arms (GPU library and CPU reference) are fed byte-identical frames.
 *
 * Texture: three octaves of seeded value noise (lattice cells of 32 / 12 / 5
 * px, weights .5 / .3 / .2, bilinear), quantised to u8.  Frame t samples the
 * texture at  A_t * (x - cx, y - cy) + (cx, cy) + t * (vx, vy)  where A_t is a
 * rotation by t*rot_deg degrees times a scale of scale_per_frame^t (identity for
 * the pure-translation sequence).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>

static inline uint32_t lattice_hash(uint32_t seed, int octave, int ix, int iy)
{
  /* one LCG step per mixed-in word (s = s*1664525 + 1013904223), then a finaliser */
  uint32_t s = seed;
  s = s * 1664525u + 1013904223u + (uint32_t)octave;
  s = (s ^ (uint32_t)ix) * 1664525u + 1013904223u;
  s = (s ^ (uint32_t)iy) * 1664525u + 1013904223u;
  s ^= s >> 16; s *= 0x7feb352du; s ^= s >> 15; s *= 0x846ca68bu; s ^= s >> 16;
  return s;
}

static inline float lattice(uint32_t seed, int octave, int ix, int iy)
{
  return (float)(lattice_hash(seed, octave, ix, iy) >> 8) * (1.0f / 16777216.0f);
}

static inline float octave_value(uint32_t seed, int octave, float cell, float x, float y)
{
  const float fx = x / cell, fy = y / cell;
  const float flx = floorf(fx), fly = floorf(fy);
  const int ix = (int)flx, iy = (int)fly;
  const float ax = fx - flx, ay = fy - fly;
  const float v00 = lattice(seed, octave, ix, iy), v01 = lattice(seed, octave, ix + 1, iy);
  const float v10 = lattice(seed, octave, ix, iy + 1), v11 = lattice(seed, octave, ix + 1, iy + 1);
  return (1 - ax) * (1 - ay) * v00 + ax * (1 - ay) * v01 + (1 - ax) * ay * v10 + ax * ay * v11;
}

typedef struct {
  unsigned char *out;
  int ncols, nrows, y0, y1;
  unsigned seed;
  float a11, a12, a21, a22, tx, ty;
} synth_job;

static void *synth_rows(void *arg)
{
  const synth_job *j = (const synth_job *)arg;
  const float cx = 0.5f * j->ncols, cy = 0.5f * j->nrows;
  int x, y;
  for (y = j->y0; y < j->y1; y++)
    for (x = 0; x < j->ncols; x++) {
      const float dx = x - cx, dy = y - cy;
      const float sx = j->a11 * dx + j->a12 * dy + cx + j->tx;
      const float sy = j->a21 * dx + j->a22 * dy + cy + j->ty;
      float v = 0.5f * octave_value(j->seed, 0, 32.0f, sx, sy) +
                0.3f * octave_value(j->seed, 1, 12.0f, sx, sy) +
                0.2f * octave_value(j->seed, 2, 5.0f, sx, sy);
      v = v * 255.0f + 0.5f;
      j->out[(long)y * j->ncols + x] = (unsigned char)(v < 0.0f ? 0.0f : (v > 255.0f ? 255.0f : v));
    }
  return 0;
}

/* nthreads <= 1: run on the calling thread */
void klt_synth_frame(unsigned char *out, int ncols, int nrows, unsigned seed, float t,
                     float vx, float vy, float rot_deg, float scale_per_frame, int nthreads)
{
  const float ang = rot_deg * t * 0.017453292519943295f;
  const float sc = powf(scale_per_frame, t);
  synth_job jobs[64];
  pthread_t th[64];
  int k;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 64) nthreads = 64;
  if (nthreads > nrows) nthreads = nrows > 0 ? nrows : 1;
  for (k = 0; k < nthreads; k++) {
    synth_job *j = &jobs[k];
    j->out = out; j->ncols = ncols; j->nrows = nrows; j->seed = seed;
    j->y0 = (int)((long)nrows * k / nthreads);
    j->y1 = (int)((long)nrows * (k + 1) / nthreads);
    j->a11 = sc * cosf(ang); j->a12 = -sc * sinf(ang);
    j->a21 = sc * sinf(ang); j->a22 = sc * cosf(ang);
    j->tx = t * vx; j->ty = t * vy;
  }
  for (k = 1; k < nthreads; k++) pthread_create(&th[k], 0, synth_rows, &jobs[k]);
  synth_rows(&jobs[0]);
  for (k = 1; k < nthreads; k++) pthread_join(th[k], 0);
}
