for nb in 128 256 512; do for N in 65521 33554393; do GFFM_ELIM_NB0=$nb python tools/pluq_once.py 16384 $N 2>&1 | tail -1 | sed "s/^/NB0=$nb N=$N: /"; done; done
