for pf in 0 -1; do echo "== fft_prefetch=$pf"; B2F_FFT_PREFETCH=$pf python tools/fft_sweep.py --logn-min 10 --logn-max 14 2>&1 | grep "^{" ; done
