"""B200-native view-synthesis loss (placeholder, filled in below)."""
