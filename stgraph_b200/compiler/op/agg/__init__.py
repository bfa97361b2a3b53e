from .agg_op import agg_max, agg_mean, agg_min, agg_sum

__all__ = ["agg_sum", "agg_max", "agg_min", "agg_mean"]
