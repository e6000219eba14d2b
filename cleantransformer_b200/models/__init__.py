"""Model mirrors: modeling_bloom / modeling_gpt / modeling_bert (same class & parameter names as
CleanTransformer/models/*)."""
