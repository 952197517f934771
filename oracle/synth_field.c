/*
 * synth_field.c -- HOST generator of the synthetic benchmark fields of SURVEY.md 8(d) (C3 fBm, C4 gyroid, C5 sphere union).
 * TEST / BENCH INFRASTRUCTURE ONLY, part of libmc_oracle.so: it lets `bench.py --impl reference` (and tests without a GPU)
 * produce a workload's lattice without loading the product library.  Same parameter derivation (splitmix64, seeds) and the
 * same formulas as the device generator (isomc_synth_field); host and device sinf/cosf may differ in the last bits, so a
 * field made here is statistically, not bitwise, the device's -- parity tests always hand the SAME bytes to both sides.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <unistd.h>

static uint64_t splitmix64(uint64_t *state) {
    uint64_t z = (*state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static float unit24(uint64_t *state) { return (float)(splitmix64(state) >> 40) * (1.0f / 16777216.0f); }

typedef struct {
    int kind;
    uint32_t size, z_first;
    int64_t row0, row1;
    float *out;
    const float *amp, *freq, *dx, *dy, *dz, *ph, *cx, *cy, *cz, *r;
} synth_job;

static void *synth_rows(void *arg) {
    const synth_job *j = (const synth_job *)arg;
    const uint32_t size = j->size;
    const float inv = 1.0f / (float)(size - 1);
    for (int64_t row = j->row0; row < j->row1; ++row) {
        const uint32_t z = j->z_first + (uint32_t)(row / size), y = (uint32_t)(row % size);
        const float py = (float)y * inv, pz = (float)z * inv;
        float *o = j->out + (uint64_t)row * size;
        for (uint32_t x = 0; x < size; ++x) {
            const float px = (float)x * inv;
            float f = 0.0f;
            if (j->kind == 1) {
                for (int w = 0; w < 20; ++w) f += j->amp[w] * sinf(j->freq[w] * (j->dx[w] * px + j->dy[w] * py + j->dz[w] * pz) + j->ph[w]);
            } else if (j->kind == 2) {
                const float k = 6.283185307179586f * 8.0f;
                const float X = k * px, Y = k * py, Z = k * pz;
                f = sinf(X) * cosf(Y) + sinf(Y) * cosf(Z) + sinf(Z) * cosf(X);
            } else {
                f = 1e30f;
                for (int s = 0; s < 64; ++s) {
                    const float ax = px - j->cx[s], ay = py - j->cy[s], az = pz - j->cz[s];
                    f = fminf(f, sqrtf(ax * ax + ay * ay + az * az) - j->r[s]);
                }
            }
            o[x] = f;
        }
    }
    return 0;
}

/* kind: 1 fBm, 2 gyroid, 3 union of 64 spheres; fills sample layers [z_first, z_first + n_layers) of the size^2 x (size+1) lattice
 * (all host cores: generating the input is not part of any timed region) */
int oracle_synth_field(int kind, uint32_t size, uint64_t seed, uint32_t z_first, uint32_t n_layers, float *out) {
    float amp[20], freq[20], dx[20], dy[20], dz[20], ph[20], cx[64], cy[64], cz[64], r[64];
    uint64_t st = seed;
    if (size < 2 || !out || n_layers == 0) return -1;
    if (kind == 1) {
        for (int o = 0; o < 5; ++o)
            for (int k = 0; k < 4; ++k) {
                float x, y, z, len;
                do {
                    x = 2.0f * unit24(&st) - 1.0f; y = 2.0f * unit24(&st) - 1.0f; z = 2.0f * unit24(&st) - 1.0f;
                    len = sqrtf(x * x + y * y + z * z);
                } while (len < 1e-3f);
                const int w = o * 4 + k;
                dx[w] = x / len; dy[w] = y / len; dz[w] = z / len;
                ph[w] = 6.283185307179586f * unit24(&st);
                amp[w] = 1.0f / (float)(1 << o);
                freq[w] = 6.283185307179586f * 4.0f * (float)(1 << o) * ((float)(size - 1) / 511.0f);
            }
    } else if (kind == 3) {
        for (int s = 0; s < 64; ++s) {
            cx[s] = 0.1f + 0.8f * unit24(&st); cy[s] = 0.1f + 0.8f * unit24(&st); cz[s] = 0.1f + 0.8f * unit24(&st);
            r[s] = 0.05f + 0.1f * unit24(&st);
        }
    } else if (kind != 2) {
        return -1;
    }
    const int64_t rows = (int64_t)n_layers * size;
    long nt = sysconf(_SC_NPROCESSORS_ONLN);
    if (nt < 1) nt = 1;
    if (nt > 64) nt = 64;
    if (nt > rows) nt = (long)rows;
    pthread_t th[64];
    synth_job jobs[64];
    for (long t = 0; t < nt; ++t) {
        synth_job *j = &jobs[t];
        j->kind = kind; j->size = size; j->z_first = z_first; j->out = out;
        j->row0 = rows * t / nt; j->row1 = rows * (t + 1) / nt;
        j->amp = amp; j->freq = freq; j->dx = dx; j->dy = dy; j->dz = dz; j->ph = ph; j->cx = cx; j->cy = cy; j->cz = cz; j->r = r;
        if (pthread_create(&th[t], 0, synth_rows, j) != 0) { synth_rows(j); th[t] = 0; }
    }
    for (long t = 0; t < nt; ++t)
        if (th[t]) pthread_join(th[t], 0);
    return 0;
}
