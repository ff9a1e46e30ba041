"""dfcsa: B200-native DFC-SA-Res-Block hot path (host side, mirrors the reference Python API)."""
