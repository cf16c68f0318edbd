// placeholder until the Quatro oracle lands
