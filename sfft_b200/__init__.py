"""sfft_b200 -- B200-native sparse FFT behind the sfft.h C API.

`sfft_b200.sfft` is the module mirroring the reference's Python binding
(python/sfft/sfft.py): `import sfft_b200.sfft as sfft; sfft.sfft(n, k, version)`.
`sfft_b200.build` compiles the in-tree libsfft.so for sm_100a.
"""
