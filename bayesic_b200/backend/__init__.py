"""Binding of the host-side algebra to the sm_100a executor (C-ABI via ctypes)."""
