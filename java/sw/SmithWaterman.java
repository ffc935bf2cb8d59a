package sw ;

import org.apache.spark.api.java.function.Function3 ;

import scala.Tuple2 ;

import java.nio.charset.StandardCharsets ;
import java.util.ArrayList ;
import java.util.Collections ;
import java.util.List ;

/**
 * Drop-in host layer for the reference's sw.SmithWaterman (Java 1.8): same package, class, nested public
 * class and method signature (reference SmithWaterman.java:35, 62); the body marshals to libswb200
 * (hand-written sm_100a CUDA) through the JNI shim and rebuilds the reference's object graph
 * {@code Tuple2<maxScore, ArrayList<Tuple2<beginning, {refAln, readAln}>>>} (SmithWaterman.java:91).
 * Written fresh against include/swb200.h; NOT COMPILED in the build environment (no JDK there).
 *
 * Per-pair calls (the unchanged driver) go through the native submission queue, which coalesces the
 * calls of concurrent task threads; the throughput path is {@link #alignAll}, which the batched MapRef
 * of java/sw/Distribution.java calls once per reference file (all references x all reads in one launch
 * sequence).
 */
@SuppressWarnings( "serial" )
public class SmithWaterman
{
	public static class OptAlignments implements Function3< String[] , int[] , char[] , Tuple2<Integer,ArrayList<Tuple2<Integer,String[]>>> >
	{
		/** seqs = {reference, read}; alignScores = {match, mismatch, gap}; alignTypes never leave the operator. */
		@Override
		public Tuple2<Integer,ArrayList<Tuple2<Integer,String[]>>> call( String[] seqs , int[] alignScores , char[] alignTypes )
		{
			// one request to the native submission queue: the calls of concurrent Spark task threads (the unchanged
			// MapRef, reference Distribution.java:419-426) are coalesced into one launch sequence per batch
			long res = NativeSW.alignPair( NativeSW.context() , seqs[0].getBytes( StandardCharsets.ISO_8859_1 ) ,
					seqs[1].getBytes( StandardCharsets.ISO_8859_1 ) , alignScores[0] , alignScores[1] , alignScores[2] , 0 ) ;
			try { return unmarshal( res , Collections.singletonList(seqs[0]) , Collections.singletonList(seqs[1]) ).get(0).get(0) ; }
			finally { NativeSW.resultFree( res ) ; }
		}
	}

	/** result.get(ref).get(read) = what OptAlignments.call returns for that pair. */
	public static List<List<Tuple2<Integer,ArrayList<Tuple2<Integer,String[]>>>>> alignAll( List<String> refs , List<String> reads , int[] alignScores , int flags )
	{
		long ctx = NativeSW.context() ;
		long[] refOff = new long[refs.size() + 1] , readOff = new long[reads.size() + 1] ;
		byte[] refBytes = NativeSW.pack( refs , refOff ) ;
		byte[] readBytes = NativeSW.pack( reads , readOff ) ;
		long refset = NativeSW.refsetLoad( ctx , refBytes , refOff ) ;
		try
		{
			long res = NativeSW.align( ctx , refset , readBytes , readOff , alignScores[0] , alignScores[1] , alignScores[2] , flags ) ;
			try { return unmarshal( res , refs , reads ) ; }
			finally { NativeSW.resultFree( res ) ; }
		}
		finally { NativeSW.refsetFree( refset ) ; }
	}

	/** flat result arrays -> the reference's nested tuples; max cells in the reference's list order */
	static List<List<Tuple2<Integer,ArrayList<Tuple2<Integer,String[]>>>>> unmarshal( long res , List<String> refs , List<String> reads )
	{
		int nReads = reads.size() ;
		int[] scores = NativeSW.scores( res ) ;
		long[] offs = NativeSW.cellOffsets( res ) ;
		int[] cells = NativeSW.cells( res ) , begs = NativeSW.beginnings( res ) , lens = NativeSW.opLens( res ) ;
		List<List<Tuple2<Integer,ArrayList<Tuple2<Integer,String[]>>>>> all = new ArrayList<List<Tuple2<Integer,ArrayList<Tuple2<Integer,String[]>>>>>( refs.size() ) ;
		byte[][] readBytes = new byte[nReads][] ;
		for( int q = 0 ; q < nReads ; q++ ) readBytes[q] = reads.get(q).getBytes( StandardCharsets.ISO_8859_1 ) ;
		for( int r = 0 ; r < refs.size() ; r++ )
		{
			byte[] refBytes = refs.get(r).getBytes( StandardCharsets.ISO_8859_1 ) ;
			List<Tuple2<Integer,ArrayList<Tuple2<Integer,String[]>>>> row = new ArrayList<Tuple2<Integer,ArrayList<Tuple2<Integer,String[]>>>>( nReads ) ;
			for( int q = 0 ; q < nReads ; q++ )
			{
				int pair = r * nReads + q ;
				int score = scores[pair] ;
				ArrayList<Tuple2<Integer,String[]>> opt = new ArrayList<Tuple2<Integer,String[]>>() ;
				if( score == 0 )
				{
					// every cell ties at 0: m*n entries with beginning 0 and empty strings (SmithWaterman.java:180-185, :380)
					long count = NativeSW.pairCellCount( res , pair ) ;
					for( long k = 0 ; k < count ; k++ ) opt.add( new Tuple2<Integer,String[]>( 0 , new String[]{ "" , "" } ) ) ;
				}
				else for( long c = offs[pair] ; c < offs[pair + 1] ; c++ )
				{
					byte[][] aln = NativeSW.materialize( res , c , lens[(int) c] , refBytes , readBytes[q] ) ;
					String[] strs = { new String( aln[0] , StandardCharsets.ISO_8859_1 ) , new String( aln[1] , StandardCharsets.ISO_8859_1 ) } ;
					opt.add( new Tuple2<Integer,String[]>( begs[(int) c] , strs ) ) ;
				}
				row.add( new Tuple2<Integer,ArrayList<Tuple2<Integer,String[]>>>( score , opt ) ) ;
			}
			all.add( row ) ;
		}
		return all ;
	}
}
