"""Drop-in replacement for the reference package `shot_vae_model` (same import path, same class
names, same state_dict keys); the arithmetic runs in libshotvae (sm_100a CUDA)."""
