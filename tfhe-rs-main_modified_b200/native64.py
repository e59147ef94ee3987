"""native64::{Plan32, Plan52} (reference: tfhe-ntt/src/native64.rs)."""
from ._native_plan import make

Plan32 = make(2, False, "native64::Plan32")
Plan52 = make(3, False, "native64::Plan52 (always available here; the reference needs AVX512-IFMA)")
