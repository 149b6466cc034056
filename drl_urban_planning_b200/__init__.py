"""B200-native PPO-update path for DRL-urban-planning."""
__version__ = "0.1.0"
