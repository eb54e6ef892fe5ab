"""B200-native (sm_100a) SD-2 UNet noise-prediction training step behind the reference's model API."""
__version__ = '0.1.0'
