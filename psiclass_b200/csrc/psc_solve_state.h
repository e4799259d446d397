// psc_solve_state.h -- host-side buffers of psc_solve (included by psc_api.cu after DevBuf)
#pragma once
// device buffers of one pass of the per-unit kernels (pick / refine / EM / emit).  Two sets: the pass over
// the enumerated components runs on its own stream while the DP kernel is still busy on the main one.
struct UnitBufs {
	DevBuf scr_off, scratch, pred_off, pred, n_pred, n_raw, n_exon, order, tx_off, ex_off, dbg;
	DevBuf o_comp, o_sample, o_chr, o_gene, o_strand, o_exon_off, o_exon_start, o_exon_end, o_fpkm, o_cov;
	HostBuf stage; // pinned staging of the results
	void release()
	{
		stage.release();
		DevBuf *all[] = {&scr_off, &scratch, &pred_off, &pred, &n_pred, &n_raw, &n_exon, &order, &tx_off, &ex_off, &dbg, &o_comp,
		                 &o_sample, &o_chr, &o_gene, &o_strand, &o_exon_off, &o_exon_start, &o_exon_end, &o_fpkm, &o_cov};
		for (DevBuf *b : all) b->release();
	}
};

// device buffers of one DP launch sequence (k_dp_units + k_dp_merge)
struct DpBufs {
	DevBuf list, dp_scr_off, dp_order, dp_tx, dp_ntx, dp_units, dp_arena, dp_ovf, dp_merge_scr, dp_packets;
	DevBuf *scratch = nullptr;
	void release()
	{
		DevBuf *all[] = {&list, &dp_scr_off, &dp_order, &dp_tx, &dp_ntx, &dp_units, &dp_arena, &dp_ovf, &dp_merge_scr, &dp_packets};
		for (DevBuf *b : all) b->release();
	}
};

// transcripts copied back to the host, in the order the passes produced them
struct ResultSet {
	std::vector<int> r_comp, r_sample, r_chr, r_gene, r_exon_start, r_exon_end;
	std::vector<int8_t> r_strand;
	std::vector<long long> r_exon_off;
	std::vector<double> r_fpkm, r_cov;
	void reset()
	{
		r_comp.clear(); r_sample.clear(); r_chr.clear(); r_gene.clear(); r_strand.clear();
		r_exon_off.assign(1, 0); r_exon_start.clear(); r_exon_end.clear(); r_fpkm.clear(); r_cov.clear();
	}
};

struct SolveState : ResultSet {
	DevBuf can_start, can_end, edge_support, cs00, at_cnt, use_dp, def_gene, n_cand, cand_row_off, cand_bits_off,
	    cand_bits, cand_first, cand_last, cand_corr, cand_sup1, cand_sup2, error_flag, fscratch, fs_off, list, stk, dp_dbg,
	    dp_heavy_scratch, dp_rest_scratch;
	UnitBufs ub[3]; // heavy chain, bulk DP chain, enumerated-components chain
	// db[0]: the heavy DP components; db[1]: the others.  The scratch buffers are only used by the walk-engine fallback.
	DpBufs db[2];
	SolveState()
	{
		db[0].scratch = &dp_heavy_scratch;
		db[1].scratch = &dp_rest_scratch;
	}
	std::vector<int> h_at_cnt, h_use_dp, h_n_cand;
	ResultSet passR, passH, passE, passRE; // per-pass results (kept across calls: no reallocation, no page faults after the first step)
	// results, host side (filled chunk by chunk): the ResultSet base
	void release()
	{
		DevBuf *all[] = {&can_start, &can_end, &edge_support, &cs00, &at_cnt, &use_dp, &def_gene, &n_cand,
		                 &cand_row_off, &cand_bits_off, &cand_bits, &cand_first, &cand_last, &cand_corr, &cand_sup1,
		                 &cand_sup2, &error_flag, &fscratch, &fs_off, &list, &stk, &dp_dbg, &dp_heavy_scratch, &dp_rest_scratch};
		for (DevBuf *b : all) b->release();
		ub[0].release();
		ub[1].release();
		ub[2].release();
		db[0].release();
		db[1].release();
	}
};
