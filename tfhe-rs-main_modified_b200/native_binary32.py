"""native_binary32::{Plan32, Plan52} (reference: tfhe-ntt/src/native_binary32.rs)."""
from ._native_plan import make

Plan32 = make(5, True, "native_binary32::Plan32")
Plan52 = make(6, True, "native_binary32::Plan52 (always available here; the reference needs AVX512-IFMA)")
