/* TEST INFRASTRUCTURE ONLY. Toy cuRAND device API: VSIDS.cuh includes it although USE_VSIDS is off. */
#ifndef GPSAT_ORACLE_CURAND_STUB_H
#define GPSAT_ORACLE_CURAND_STUB_H
struct curandState { unsigned long long s; };
typedef curandState curandState_t;
static inline void curand_init(unsigned long long seed, unsigned long long seq, unsigned long long, curandState *st)
{ st->s = seed * 6364136223846793005ULL + seq + 1442695040888963407ULL; }
static inline unsigned curand(curandState *st)
{ st->s = st->s * 6364136223846793005ULL + 1442695040888963407ULL; return (unsigned)(st->s >> 33); }
static inline float curand_uniform(curandState *st) { return (curand(st) + 1.0f) / 2147483649.0f; }
#endif
