"""damar_b200 -- B200-native daligner overlap kernel behind DAmar's filter.h interface."""
__version__ = "0.1.0"
