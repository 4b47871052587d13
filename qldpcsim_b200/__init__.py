"""qldpcsim_b200 -- B200-native (sm_100a) batched quantum-LDPC decoder, drop-in for the decoder path of
albertogp71/qLDPCsim (`decoders.py` as driven by `simulator.simulate()`).

    from qldpcsim_b200 import decoders, simulator, pcmlibrary

Sub-modules are imported lazily so that host-only helpers (pcm, pcmlibrary, sampler, bitpack) work without CUDA.
"""
__version__ = "0.1.0"
__all__ = ["decoders", "simulator", "pcm", "pcmlibrary", "sampler", "bitpack", "build"]
