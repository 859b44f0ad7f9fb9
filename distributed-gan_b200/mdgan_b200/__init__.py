"""B200-native engine behind the MD-GAN actor API."""
