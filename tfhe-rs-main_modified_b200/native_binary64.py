"""native_binary64::{Plan32, Plan52} (reference: tfhe-ntt/src/native_binary64.rs)."""
from ._native_plan import make

Plan32 = make(7, True, "native_binary64::Plan32")
Plan52 = make(8, True, "native_binary64::Plan52 (always available here; the reference needs AVX512-IFMA)")
