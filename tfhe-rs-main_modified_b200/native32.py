"""native32::{Plan32, Plan52} (reference: tfhe-ntt/src/native32.rs)."""
from ._native_plan import make

Plan32 = make(0, False, "native32::Plan32")
Plan52 = make(1, False, "native32::Plan52 (always available here; the reference needs AVX512-IFMA)")
