"""Tacotron-2 decoder path (secondary hot path): checkpoint reader, host-side mirror, CUDA decoder step."""
