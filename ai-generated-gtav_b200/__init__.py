"""gtav_b200: B200-native inference hot path of AI-Generated-GTAV behind the reference's Python surface.

    from gtav_b200.model.dit import DiT_models          # reference model/dit.py
    from gtav_b200.model.vae import VAE_models          # reference model/vae.py
    from gtav_b200.train_dit import denoise_step        # reference train_dit.py:30
    from gtav_b200.utils import sigmoid_beta_schedule   # reference utils.py:30
    from gtav_b200.sampler import Sampler               # graph-captured generate.py:186-244 loop

(The directory can also be put on sys.path directly, in which case the reference's own import lines
`from model.dit import DiT_models` etc. resolve to these modules unchanged.)
"""
