"""Stand-in for fairscale.nn.checkpoint.checkpoint_wrapper (models/lemevit.py:19; only used with use_checkpoint_stages).  TEST INFRASTRUCTURE."""
