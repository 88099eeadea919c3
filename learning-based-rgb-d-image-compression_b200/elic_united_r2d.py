"""ELIC_united_R2D — unidirectional RGB->depth variant (reference models/elic_united_R2D.py:9-326).

The rgb branch is a plain single-modality ELIC chain (its transforms, hyper-synthesis and
context model never see depth); the depth branch is conditioned on rgb exactly as in ELIC_united.
"""
from .elic_united import ELIC_united


class ELIC_united_R2D(ELIC_united):
    cross = False
