from .torchutils import (cbrt, random_orthogonal, create_alternating_binary_mask, create_mid_split_binary_mask, create_random_binary_mask,
                         get_num_parameters, logabsdet, merge_leading_dims, repeat_rows, searchsorted,
                         split_leading_dim, sum_except_batch, tensor2numpy, tile)
from .typechecks import is_bool, is_int, is_nonnegative_int, is_positive_int, is_power_of_two
