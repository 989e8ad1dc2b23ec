"""facialmmt_b200: B200-native (sm_100a) inference forward path of NUSTM/FacialMMT behind the reference's
src/models.py API. Hand-written CUDA in csrc/, reached only through the C ABI in include/facialmmt_b200.h."""
__version__ = "0.1.0"
