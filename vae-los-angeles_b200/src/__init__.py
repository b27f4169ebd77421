"""Drop-in replacement of the reference's `src` package (marcin119a/vae-los-angeles) on B200.

Same import surface (reference src/models/__init__.py:4-8, src/utils/__init__.py:4, src/data/__init__.py:4,
src/config.py:7); the arithmetic runs in libvla_b200 (sm_100a).  Put this directory's parent first on
PYTHONPATH and the reference's train_rna2dna.py / train_dna2rna.py run unmodified.
"""
__version__ = "1.0.0"
