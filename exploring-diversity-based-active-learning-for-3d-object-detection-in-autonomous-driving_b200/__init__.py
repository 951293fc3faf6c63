"""B200-native diversity-based active selection (drop-in for the reference det3d.selectors hot path)."""
