"""Field components of the K-Planes path: spatial distortion and the appearance embedding."""
