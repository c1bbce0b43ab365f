"""kronmult993_b200 -- B200-native ``kronmult_batched``."""
