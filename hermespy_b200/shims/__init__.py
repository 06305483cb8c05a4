"""Import shims that let an unmodified HermesPy run where its cluster / plotting dependencies are absent."""
