package sw ;

import org.apache.spark.api.java.function.Function3 ;

import scala.Tuple2 ;

import java.lang.foreign.Arena ;
import java.lang.foreign.MemorySegment ;
import java.util.ArrayList ;
import java.util.Collections ;
import java.util.List ;

import static java.lang.foreign.ValueLayout.ADDRESS ;
import static java.lang.foreign.ValueLayout.JAVA_INT ;
import static java.lang.foreign.ValueLayout.JAVA_LONG ;

/**
 * Drop-in host layer for the reference's sw.SmithWaterman: same package, class, nested
 * public class and method signature; the body marshals to libswb200 (sm_100a CUDA) and
 * unmarshals the result into the reference's object graph.  Written fresh against
 * include/swb200.h; NOT COMPILED in the build environment (no JDK there).
 *
 * Per-pair calls work but pay one native round trip each; the throughput path is
 * {@link #alignAll}, which a MapRef / DistributeReference integration calls once per
 * reference file (INTEGRATION.md).
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
			return alignAll( Collections.singletonList(seqs[0]) , Collections.singletonList(seqs[1]) , alignScores ).get(0).get(0) ;
		}
	}

	/** result.get(ref).get(read) = what OptAlignments.call returns for that pair. */
	public static List<List<Tuple2<Integer,ArrayList<Tuple2<Integer,String[]>>>>> alignAll( List<String> refs , List<String> reads , int[] alignScores )
	{
		try( Arena a = Arena.ofConfined() )
		{
			MemorySegment[] r = NativeSW.pack( a , refs ) ;
			MemorySegment[] q = NativeSW.pack( a , reads ) ;
			MemorySegment out = a.allocate( ADDRESS ) ;
			NativeSW.check( (int) NativeSW.REFSET_LOAD.invokeExact( NativeSW.CTX , (long) refs.size() , r[0] , r[1] , out ) ) ;
			MemorySegment refset = out.get( ADDRESS , 0 ) ;
			try
			{
				NativeSW.check( (int) NativeSW.ALIGN.invokeExact( NativeSW.CTX , refset , (long) reads.size() , q[0] , q[1] ,
						alignScores[0] , alignScores[1] , alignScores[2] , 0 , out ) ) ;
				MemorySegment res = out.get( ADDRESS , 0 ) ;
				try { return unmarshal( a , res , refs , reads ) ; }
				finally { NativeSW.RESULT_FREE.invokeExact( res ) ; }
			}
			finally { NativeSW.REFSET_FREE.invokeExact( refset ) ; }
		}
		catch( RuntimeException e ) { throw e ; }
		catch( Throwable t ) { throw new RuntimeException( t ) ; }
	}

	private static List<List<Tuple2<Integer,ArrayList<Tuple2<Integer,String[]>>>>> unmarshal( Arena a , MemorySegment res , List<String> refs , List<String> reads ) throws Throwable
	{
		long nReads = reads.size() ;
		MemorySegment scores = ((MemorySegment) NativeSW.SCORES.invokeExact( res )).reinterpret( 4L * refs.size() * nReads ) ;
		MemorySegment offs = ((MemorySegment) NativeSW.CELL_OFFS.invokeExact( res )).reinterpret( 8L * (refs.size() * nReads + 1) ) ;
		MemorySegment pi = a.allocate( JAVA_INT ) , pj = a.allocate( JAVA_INT ) , pb = a.allocate( JAVA_INT ) , pl = a.allocate( JAVA_INT ) ;
		List<List<Tuple2<Integer,ArrayList<Tuple2<Integer,String[]>>>>> all = new ArrayList<>( refs.size() ) ;
		for( int ref = 0 ; ref < refs.size() ; ref++ )
		{
			List<Tuple2<Integer,ArrayList<Tuple2<Integer,String[]>>>> row = new ArrayList<>( reads.size() ) ;
			MemorySegment refBytes = a.allocateFrom( refs.get(ref) , java.nio.charset.StandardCharsets.ISO_8859_1 ) ;
			for( int rd = 0 ; rd < reads.size() ; rd++ )
			{
				long pair = ref * nReads + rd ;
				int score = scores.getAtIndex( JAVA_INT , pair ) ;
				long count = (long) NativeSW.CELL_COUNT.invokeExact( res , pair ) ;      // m*n when score == 0
				long base = offs.getAtIndex( JAVA_LONG , pair ) ;
				ArrayList<Tuple2<Integer,String[]>> opt = new ArrayList<>( (int) count ) ;
				MemorySegment readBytes = a.allocateFrom( reads.get(rd) , java.nio.charset.StandardCharsets.ISO_8859_1 ) ;
				for( long k = 0 ; k < count ; k++ )
				{
					NativeSW.check( (int) NativeSW.PAIR_CELL.invokeExact( res , pair , k , pi , pj , pb , pl ) ) ;
					int len = pl.get( JAVA_INT , 0 ) ;
					String[] aligned = { "" , "" } ;
					if( score != 0 )
					{
						MemorySegment ra = a.allocate( len + 1L ) , qa = a.allocate( len + 1L ) ;
						NativeSW.check( (int) NativeSW.MATERIALIZE.invokeExact( res , base + k , refBytes , (long) refs.get(ref).length() ,
								readBytes , (long) reads.get(rd).length() , ra , qa , len + 1L ) ) ;
						aligned[0] = ra.getString( 0 , java.nio.charset.StandardCharsets.ISO_8859_1 ) ;
						aligned[1] = qa.getString( 0 , java.nio.charset.StandardCharsets.ISO_8859_1 ) ;
					}
					opt.add( new Tuple2<Integer,String[]>( pb.get(JAVA_INT,0) , aligned ) ) ;
				}
				row.add( new Tuple2<Integer,ArrayList<Tuple2<Integer,String[]>>>( score , opt ) ) ;
			}
			all.add( row ) ;
		}
		return all ;
	}
}
