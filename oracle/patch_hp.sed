# Applied at build time to a COPY of high-precision-anticyclic-fft/src/code.cpp (SURVEY.md 8c: the file needs NTL, which is absent,
# and includes <bmi2intrin.h> directly, which gcc 13 rejects).  Line-addressed; the Makefile greps for the markers.
# everything after the standard includes goes into a namespace: the harness links this next to the circuit-bootstrapping PoC
8s/$/\nnamespace hpref { \/\/ ORACLE_PATCH_HP_NS/
# :139 the BMI2 intrinsics come from <immintrin.h> (force-included on the command line, outside the namespace)
139s/.*/\/\/ ORACLE_PATCH_HP_BMI2/
# :241-277 NTL twiddle generator -> the same two functions on libquadmath (hp_twiddles_quadmath.inc)
241,277d
# the harness owns main()
516s/int main(/int hp_original_main(/
