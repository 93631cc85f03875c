"""Autograd operators over the C ABI (include/scp_b200.h)."""
