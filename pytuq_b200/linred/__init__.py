"""KL expansion and KL + PC surrogate: the multi-output end of the PC hot path."""
