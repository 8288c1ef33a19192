# Applied at build time to a COPY of cb/poc_CircuitBootstrapping.cpp (SURVEY.md Appendix B).  Line-addressed on purpose:
# the reference tree is immutable, and the Makefile greps for the markers so a silent mismatch fails the build.
# D3 (:542) rotate the test vector by 2N - bbar, as cb/lwe_functions.cpp:385 does
542s/const int bbar = abar\[n_lvl0\];/const int bbar = (2*N_lvl2 - abar[n_lvl0]) % (2*N_lvl2); \/\/ ORACLE_PATCH_D3/
# D2 (:593,:596,:597) signs and indices of (X^a - 1), as cb/numeric_functions.cpp:311-322
593s/= acc1->a\[q\]\.coefs\[j-aibar+N_lvl2\]/= -acc1->a[q].coefs[j-aibar+N_lvl2]/
593s/$/ \/\/ ORACLE_PATCH_D2/
596s/coefs\[j-aibar+N_lvl2\]/coefs[j-aibar+2*N_lvl2]/
597s/coefs\[j-aibar\]/coefs[j-aibar+N_lvl2]/
# D1 (:618) use bootstrapping-key entry i, as cb/lwe_functions.cpp:352 passes bkFFT+i
618s/&bkFFT->allsamples\[p\]/\&bkFFT[i].allsamples[p]/
# the harness owns main()
912s/int main(/int poc_original_main(/
