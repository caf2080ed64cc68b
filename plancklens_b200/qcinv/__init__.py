"""Conjugate-gradient inverse-variance filter of plancklens (reference: plancklens/qcinv/), B200-resident.

Vectors live on the GPU (`util_alm.dalm`, `util_alm.eblm`); every operator application runs through the CUDA
transforms of libplk_b200 and the alm BLAS-1 kernels of the same library."""
