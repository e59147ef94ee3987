"""native_binary128::Plan32 (reference: tfhe-ntt/src/native_binary128.rs)."""
from ._native_plan import make

Plan32 = make(9, True, "native_binary128::Plan32")
