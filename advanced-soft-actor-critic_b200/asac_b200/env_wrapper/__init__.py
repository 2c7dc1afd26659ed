"""Environment-side interface types that plugin packages import next to their ``nn`` files
(reference: algorithm/env_wrapper/env_wrapper.py, obs_preprocessor_wrapper.py).  Only the abstract
interface lives here — concrete environments (Unity, gym, offline datasets) are outside the B200 path."""
