// psc_solve_state.h -- host-side buffers of psc_solve (included by psc_api.cu after DevBuf)
#pragma once
struct SolveState {
	DevBuf can_start, can_end, edge_support, cs00, at_cnt, use_dp, def_gene, n_cand, cand_row_off, cand_bits_off,
	    cand_bits, cand_first, cand_last, cand_corr, cand_sup1, cand_sup2, error_flag, fscratch, fs_off, list, scr_off,
	    scratch, pred_off, pred, n_pred, n_raw, n_exon, order, dp_order, tx_off, ex_off, stk, dp_tx, dp_ntx, dp_units, dp_dbg;
	DevBuf o_comp, o_sample, o_chr, o_gene, o_strand, o_exon_off, o_exon_start, o_exon_end, o_fpkm, o_cov;
	std::vector<int> h_at_cnt, h_use_dp, h_n_cand;
	// results, host side (filled chunk by chunk)
	std::vector<int> r_comp, r_sample, r_chr, r_gene, r_exon_start, r_exon_end;
	std::vector<int8_t> r_strand;
	std::vector<long long> r_exon_off;
	std::vector<double> r_fpkm, r_cov;
	void release()
	{
		DevBuf *all[] = {&can_start, &can_end, &edge_support, &cs00, &at_cnt, &use_dp, &def_gene, &n_cand,
		                 &cand_row_off, &cand_bits_off, &cand_bits, &cand_first, &cand_last, &cand_corr, &cand_sup1,
		                 &cand_sup2, &error_flag, &fscratch, &fs_off, &list, &scr_off, &scratch, &pred_off, &pred,
		                 &n_pred, &n_raw, &n_exon, &order, &dp_order, &tx_off, &ex_off, &stk, &dp_tx, &dp_ntx, &dp_units, &dp_dbg, &o_comp, &o_sample,
		                 &o_chr, &o_gene, &o_strand, &o_exon_off, &o_exon_start, &o_exon_end, &o_fpkm, &o_cov};
		for (DevBuf *b : all) b->release();
	}
};

