/*---------------------------------------------------------------------------------------------
 *  NEW FILE for src/vs/workbench/contrib/senweaver/common/apoScoringService.ts
 *
 *  Common-layer contract + renderer-side proxy of the B200 scoring engine.  The `common`
 *  layer may not load native code (eslint.config.js:94-111), so this file only declares the
 *  service and builds a ProxyChannel client; the implementation lives in the main process
 *  (apoScoringMainService.ts), exactly like the metrics service of the reference
 *  (common/metricsService.ts:25-50 / electron-main/metricsMainService.ts:35).
 *
 *  This file cannot be compiled in the engine's build image (no tsc / node); it is the
 *  reference-side binding INTEGRATION.md describes.
 *--------------------------------------------------------------------------------------------*/
import { VSBuffer } from '../../../../base/common/buffer.js';
import { ProxyChannel } from '../../../../base/parts/ipc/common/ipc.js';
import { createDecorator } from '../../../../platform/instantiation/common/instantiation.js';
import { registerSingleton, InstantiationType } from '../../../../platform/instantiation/common/extensions.js';
import { IMainProcessService } from '../../../../platform/ipc/common/mainProcessService.js';

export const APO_SCORING_CHANNEL = 'senweaver-channel-apoScoring';

/** Raw result blocks; layouts are those of include/apo_b200.h (see traceRecordCodec.ts). */
export interface ApoScoreBlocks {
	scores: VSBuffer;   // float64[C]  (-Infinity: candidate without a non-null evaluation)
	counts: VSBuffer;   // uint64[C]
	topk: VSBuffer;     // int32[K]    score desc, ties -> lower index
	report: VSBuffer;   // struct apo_corpus_report (784 bytes)
}

/** Form T of a float32[C][T][9] tensor (include/apo_b200.h): tl uint16[C][T] | th uint8[C][T] = the 24-bit dictionary index of every
 *  evaluation; tbookPc uint32[n] / tbookPd uint16[n] = the n distinct evaluations as Form P pairs; codebook uint32[8][256], d2book
 *  uint32[4096] = the fp32 bit patterns the codes stand for. */
export interface ApoTupleBuffers {
	tl: VSBuffer; th: VSBuffer; tbookPc: VSBuffer; tbookPd: VSBuffer; codebook: VSBuffer; d2book: VSBuffer; nTuples: number;
}

/** What a resident scoring call selects (apo_score_opts of include/apo_b200.h). */
export interface ApoResidentQuery {
	C: number;              // candidates uploaded with dimsUpload / rolloutsUpload
	K: number;              // beam width (APOConfig.beamWidth, APO:288)
	source?: 0 | 1;         // 0 = dims (Form D / Form Q), 1 = per-(candidate, trace) records
	corpus?: 0 | 1;         // also build the corpus report from the uploaded corpus
	first?: number;         // window [first, first + count) of the trace axis; count 0 = to the end
	count?: number;
}

export interface IApoScoringService {
	readonly _serviceBrand: undefined;
	/** false when the addon or a B200 is unavailable: callers keep their state untouched and warn. */
	isAvailable(): Promise<boolean>;
	/** TraceCollectorService._computeRewardSignals for n packed records (apo_reward_batch).  Calls made in the same tick
	 *  are coalesced into one engine call by the main-process service. */
	rewardBatch(records: VSBuffer): Promise<{ dims: VSBuffer; masks: VSBuffer; finals: VSBuffer }>;
	/** The string stored under 'senweaver.traceCollector.data' -> packed Form R records (apo_records_from_json):
	 *  lets a stored corpus reach the engine without materialising ConversationTrace objects. */
	recordsFromJson(persisted: string): Promise<VSBuffer>;

	// ---- resident path: uploaded once, scored many times (only the result block crosses PCIe again)
	/** float32[C][T][9], NaN = dimension absent.  compact: keep the lossless 14-byte layout on the device. */
	dimsUpload(dims: VSBuffer, C: number, T: number, compact?: boolean): Promise<void>;
	/** [C][T] trace records, 32-byte apo_record or 16-byte apo_record16 rows; dims are derived on the device (TCS:668-763). */
	rolloutsUpload(records: VSBuffer, rowBytes: 32 | 16, C: number, T: number): Promise<void>;
	/** The traces _buildReport walks (APO:498-625), one 32-byte record each; idxBase = index of the first one. */
	corpusUpload(records: VSBuffer, idxBase?: number): Promise<void>;
	/** Same from the persisted JSON string; resolves to the number of traces. */
	corpusUploadJson(persisted: string, idxBase?: number): Promise<number>;
	scoreResident(query: ApoResidentQuery): Promise<ApoScoreBlocks>;

	// ---- one-shot path: host buffers streamed through the device, nothing stays resident
	/** dims: float32[C][T][9] (pass C = 1, T = 4 of NaN for report-only calls). */
	score(dims: VSBuffer, C: number, T: number, corpus: VSBuffer | undefined, K: number): Promise<ApoScoreBlocks>;
	/** [C][T] trace records (the IDE's own representation): 2.2x fewer PCIe bytes than fp32 dims with 16-byte rows. */
	scoreHostRecords(records: VSBuffer, rowBytes: 32 | 16, C: number, T: number, corpus: VSBuffer | undefined, K: number): Promise<ApoScoreBlocks>;
	/** fp32 evaluations -> Form T (include/apo_b200.h): a 24-bit dictionary index per evaluation (3 bytes instead of 36) + the
	 *  dictionary of distinct evaluations.  For evaluations that are scored more than once, kept, or moved between processes. */
	encodeTuples(dims: VSBuffer, C: number, T: number): Promise<ApoTupleBuffers | undefined>;
	scoreTuples(tuples: ApoTupleBuffers, C: number, T: number, corpus: VSBuffer | undefined, K: number): Promise<ApoScoreBlocks>;
}

export const IApoScoringService = createDecorator<IApoScoringService>('senweaverApoScoringService');

/** Renderer-side client: everything that crosses the channel is async (metricsService.ts:47). */
export class ApoScoringService implements IApoScoringService {
	readonly _serviceBrand: undefined;
	private readonly _proxy: IApoScoringService;

	constructor(@IMainProcessService mainProcessService: IMainProcessService) {
		this._proxy = ProxyChannel.toService<IApoScoringService>(mainProcessService.getChannel(APO_SCORING_CHANNEL));
	}

	isAvailable(): Promise<boolean> { return this._proxy.isAvailable(); }
	rewardBatch(records: VSBuffer) { return this._proxy.rewardBatch(records); }
	recordsFromJson(persisted: string) { return this._proxy.recordsFromJson(persisted); }
	dimsUpload(dims: VSBuffer, C: number, T: number, compact?: boolean) { return this._proxy.dimsUpload(dims, C, T, compact); }
	rolloutsUpload(records: VSBuffer, rowBytes: 32 | 16, C: number, T: number) { return this._proxy.rolloutsUpload(records, rowBytes, C, T); }
	corpusUpload(records: VSBuffer, idxBase?: number) { return this._proxy.corpusUpload(records, idxBase); }
	corpusUploadJson(persisted: string, idxBase?: number) { return this._proxy.corpusUploadJson(persisted, idxBase); }
	scoreResident(query: ApoResidentQuery) { return this._proxy.scoreResident(query); }
	score(dims: VSBuffer, C: number, T: number, corpus: VSBuffer | undefined, K: number) { return this._proxy.score(dims, C, T, corpus, K); }
	scoreHostRecords(records: VSBuffer, rowBytes: 32 | 16, C: number, T: number, corpus: VSBuffer | undefined, K: number) {
		return this._proxy.scoreHostRecords(records, rowBytes, C, T, corpus, K);
	}
	encodeTuples(dims: VSBuffer, C: number, T: number) { return this._proxy.encodeTuples(dims, C, T); }
	scoreTuples(tuples: ApoTupleBuffers, C: number, T: number, corpus: VSBuffer | undefined, K: number) { return this._proxy.scoreTuples(tuples, C, T, corpus, K); }
}

registerSingleton(IApoScoringService, ApoScoringService, InstantiationType.Delayed);
